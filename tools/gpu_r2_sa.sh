timeout 300 python tools/decode_phases.py 50 > gpurun_out/r2_decode_phases_sa.txt 2>&1; grep -E "whole|self-attention" gpurun_out/r2_decode_phases_sa.txt | cut -c1-1200
for BN in 256 128 64; do OTB_LN_BN=$BN timeout 200 python tools/encoder_time.py 2>&1 | tail -1; done
