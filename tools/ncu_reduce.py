"""Reduce an ncu report on the GPU box to text small enough to travel back (gpurun_out/ is capped at 64 MiB):
    python tools/ncu_reduce.py REPORT.ncu-rep OUT_PREFIX [--source] [--keep]
writes OUT_PREFIX.raw.csv (ncu --page raw --csv: every metric of every captured launch) and, with --source, OUT_PREFIX.src.txt:
stall-reason totals and the instructions that collected >= 0.15 % of the warp-state samples (SASS, executed count, top stall reasons),
from ncu --page source --csv.  The report itself is deleted unless --keep."""
import csv
import os
import subprocess
import sys


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    flags = sys.argv[3:]
    with open(prefix + '.raw.csv', 'w') as f:
        subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=f, check=True)
    if '--source' in flags:
        out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True, check=True).stdout
        csv.field_size_limit(10 ** 9)
        r = csv.reader(out.splitlines())
        kernel = next(r)
        hdr = next(r)
        H = {h: i for i, h in enumerate(hdr)}
        rows = [row for row in r if len(row) >= len(hdr) - 2]
        stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
        tot = sum(int(x[H['# Samples']]) for x in rows) or 1
        inst = sum(int(x[H['Instructions Executed']]) for x in rows)
        with open(prefix + '.src.txt', 'w') as f:
            f.write(f'{kernel[1] if len(kernel) > 1 else kernel}\n{len(rows)} SASS lines, {tot} warp-state samples, {inst} warp instructions executed\n')
            agg = sorted(((sum(int(x[H[s]]) for x in rows), s) for s in stalls), reverse=True)
            f.write('stall totals: ' + ', '.join(f'{s} {100 * v / tot:.1f}%' for v, s in agg[:10]) + '\n')
            by_op = {}
            for x in rows:
                op = x[H['Source']].strip().split()
                op = next((t for t in op if not t.startswith('@')), '?').split('.')[0]
                by_op[op] = by_op.get(op, 0) + int(x[H['Instructions Executed']])
            f.write('executed by opcode: ' + ', '.join(f'{k} {100 * v / max(inst, 1):.1f}%' for k, v in sorted(by_op.items(), key=lambda kv: -kv[1])[:25]) + '\n\n')
            f.write('line  samples   share  executed  SASS | top stall reasons\n')
            for i, x in enumerate(rows):
                s = int(x[H['# Samples']])
                if s < 0.0015 * tot:
                    continue
                st = sorted(((int(x[H[k]]), k) for k in stalls), reverse=True)[:3]
                f.write(f'{i:6d} {s:7d} {100 * s / tot:6.2f}% {x[H["Instructions Executed"]]:>9s}  {x[H["Source"]].strip()[:90]} | ' +
                        ', '.join(f'{k[6:]} {v}' for v, k in st if v) + '\n')
    if '--keep' not in flags:
        os.remove(rep)


if __name__ == '__main__':
    main()
