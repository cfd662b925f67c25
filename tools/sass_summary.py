"""Opcode histogram per kernel of the built library (cuobjdump -sass), as evidence of which hardware paths each kernel uses:
UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, FFMA2 = packed fp32x2 FMA, LDG/STG.E.128 = 16-byte global accesses, RED = global reductions.
    python tools/sass_summary.py > profiles/r02_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'pasta-gan-plusplus_b200', 'lib', 'libpgpp_sm100a.so')
KEYS = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'UTCBAR', 'SYNCS', 'FFMA2', 'FFMA', 'HFMA2', 'LDG.E.128', 'STG.E.128',
        'LDS.128', 'STS.128', 'RED', 'ATOMG', 'SHFL', 'BAR.SYNC', 'MUFU']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], check=True, capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
        if m and cur is not None:
            op = m.group(1)
            cur['total'] += 1
            for k in KEYS:
                if op == k or op.startswith(k + '.') or (k.count('.') and op.startswith(k)):
                    cur[k] += 1
    names = subprocess.run(['c++filt'], input='\n'.join(kernels), capture_output=True, text=True).stdout.splitlines()
    print('# SASS opcode histogram per kernel, `pasta-gan-plusplus_b200/lib/libpgpp_sm100a.so` (nvcc 12.9, sm_100a), from `cuobjdump -sass`\n')
    print('Written by `tools/sass_summary.py`. Columns: instructions of that opcode family in the kernel body (static count, not executed).\n')
    agg = collections.OrderedDict()
    for mangled, name in zip(kernels, names):
        base = re.sub(r'\(.*', '', name).replace('void ', '')
        short = re.sub(r'<.*', '', base)
        a = agg.setdefault(short, [0, collections.Counter()])
        a[0] += 1
        a[1].update(kernels[mangled])
    cols = [k for k in KEYS if any(a[1][k] for a in agg.values())]
    print('| kernel (all template instantiations summed) | inst. | SASS lines | ' + ' | '.join(cols) + ' |')
    print('|---|---|---|' + '---|' * len(cols))
    for short, (n, c) in agg.items():
        print(f'| `{short}` | {n} | {c["total"]} | ' + ' | '.join(str(c[k]) if c[k] else '' for k in cols) + ' |')
    tot = collections.Counter()
    for _, c in agg.values():
        tot.update(c)
    print(f'| **library** | {len(kernels)} | {tot["total"]} | ' + ' | '.join(str(tot[k]) if tot[k] else '' for k in cols) + ' |')
    print('\n## Tensor-core kernels, per instantiation\n')
    print('| instantiation | SASS lines | UTCHMMA | LDTM | UTMALDG | UTMASTG | UTCBAR | SYNCS | STG.E.128 | RED |')
    print('|---|---|---|---|---|---|---|---|---|---|')
    for mangled, name in zip(kernels, names):
        c = kernels[mangled]
        if c['UTCHMMA'] or c['UTMALDG']:
            print(f'| `{re.sub(r"[(].*", "", name).replace("void ", "")}` | {c["total"]} | {c["UTCHMMA"]} | {c["LDTM"]} | {c["UTMALDG"]} | {c["UTMASTG"]} | {c["UTCBAR"]} | {c["SYNCS"]} | {c["STG.E.128"]} | {c["RED"]} |')


if __name__ == '__main__':
    sys.exit(main())
