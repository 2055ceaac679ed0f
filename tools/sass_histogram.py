"""Opcode histogram of every kernel in libotb200.so (cuobjdump -sass): which kernels use the Blackwell-native paths
(UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA, UBLKCP = bulk copy) and which the legacy tensor path
(HMMA = mma.sync).  Runs without a GPU.

    python tools/sass_histogram.py > profiles/r2_sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'opentransformer_b200', 'libotb200.so')
KEYS = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR', 'HMMA', 'LDGSTS', 'LDSM', 'SYNCS', 'UCGABAR',
        'FFMA', 'MUFU', 'LDG', 'STG', 'LDS', 'STS', 'SHFL', 'BAR', 'ATOM', 'RED']


def demangle(names):
    out = subprocess.run(['c++filt'] + names, capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r'\(.*', '', o) for o in out]


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and cur:
            op = m.group(1)
            kernels[cur][op] += 1
            kernels[cur]['_total'] += 1
    names = demangle(list(kernels))
    print('cuobjdump -sass opentransformer_b200/libotb200.so : opcode counts per kernel (static instruction counts, sm_100a)')
    print('%-92s %7s  %s' % ('kernel', 'instrs', '  '.join('%s' % k for k in KEYS)))
    agg = collections.OrderedDict()
    for (mangled, c), name in zip(kernels.items(), names):
        short = re.sub(r'^void\s+', '', name)
        agg.setdefault(short, []).append(c)
    for short, cs in agg.items():
        tot = collections.Counter()
        for c in cs:
            tot.update(c)
        def cnt(k):
            return sum(v for op, v in tot.items() if op == k or op.startswith(k + '.') or (k in ('LDG', 'STG', 'LDS', 'STS', 'BAR', 'ATOM', 'RED', 'SHFL', 'MUFU', 'FFMA', 'HMMA', 'LDTM', 'STTM', 'SYNCS', 'LDSM') and op.startswith(k)))
        label = short if len(cs) == 1 else '%s  [%d instantiations]' % (short, len(cs))
        print('%-92s %7d  %s' % (label[:92], tot['_total'], '  '.join('%s=%d' % (k, cnt(k)) for k in KEYS if cnt(k))))


if __name__ == '__main__':
    sys.exit(main())
