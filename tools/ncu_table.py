"""Markdown table of the key metrics of every launch in an `ncu --page raw --csv` export:
    python tools/ncu_table.py gpurun_out/x.raw.csv [hbm_peak_GBs]"""
import csv, sys
csv.field_size_limit(10 ** 9)
rows = list(csv.reader(open(sys.argv[1])))
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6541.5
hdr = rows[0]
H = {h: i for i, h in enumerate(hdr)}
units = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
SCALE = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'nsecond': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3, 'second': 1e6}


def g(r, name, default=0.0):
    try:
        return float(r[H[name]].replace(',', ''))
    except Exception:
        return default


def gs(r, name):
    return g(r, name) * SCALE.get(units[H[name]], 1) if name in H else 0.0


print('| kernel | grid x block | regs | smem/CTA KB | time us | DRAM rd+wr MB | GB/s | % of HBM copy | dram busy % | L2 hit % | SM busy % | issue active % | achieved occ % | top stalls (warps per issue) |')
print('|---|---|---|---|---|---|---|---|---|---|---|---|---|---|')
for r in data:
    name = r[H['Kernel Name']].replace('void ', '').replace('pgpp::', '')
    name = name[:name.index('(')] if '(' in name else name
    t_us = gs(r, 'gpu__time_duration.sum')
    rd, wr = gs(r, 'dram__bytes_read.sum'), gs(r, 'dram__bytes_write.sum')
    gbs = (rd + wr) / (t_us * 1e-6) / 1e9 if t_us else 0
    stalls = {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]: g(r, k) for k in hdr
              if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio') and 'not_issued' not in k}
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
    smem = (gs(r, 'launch__shared_mem_per_block_static') + gs(r, 'launch__shared_mem_per_block_dynamic')) / 1024
    print(f"| `{name}` | {r[H['Grid Size']]} x {r[H['Block Size']]} | {g(r, 'launch__registers_per_thread'):.0f} | {smem:.1f} | "
          f"{t_us:.1f} | {(rd + wr) / 1e6:.0f} | {gbs:.0f} | {100 * gbs / peak:.0f} | {g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | {g(r, 'lts__t_sector_hit_rate.pct'):.0f} | "
          f"{g(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | {g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | {g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | "
          + ', '.join(f'{k} {v:.1f}' for k, v in top) + ' |')
