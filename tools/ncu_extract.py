#!/usr/bin/env python
"""Turn an `ncu --set full --import-source on` report into the text summaries kept under profiles/.

    python tools/ncu_extract.py gpurun_out/prof.ncu-rep [launch_index] > profiles/ncu_<kernel>_<series>_summary.txt

Needs `ncu` on PATH (it reads the report; no GPU).  Prints, for one captured launch:
  * the headline metrics (duration, DRAM bytes, executed instructions, issue utilisation, pipe activity, occupancy
    limits, shared-memory bank conflicts) and the stall reasons per issue slot;
  * the executed SASS grouped into contiguous ranges of similar execution count per warp -- the phases / loops of
    the kernel -- with each range's share of the instructions and of the stall samples and its opcode mix;
  * the instruction mix of the whole kernel.
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
]


def ncu_csv(report, page):
    out = subprocess.run(['ncu', '-i', report, '--page', page, '--csv'], capture_output=True, text=True, check=True)
    return list(csv.reader(io.StringIO(out.stdout)))


def main():
    report = sys.argv[1]
    launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    raw = ncu_csv(report, 'raw')
    hdr, units, row = raw[0], raw[1], raw[2 + launch]
    name_col = hdr.index('Kernel Name') if 'Kernel Name' in hdr else None
    if name_col is not None:
        print('kernel:', row[name_col][:160])
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            print(f'{m:66s} {row[i]:>16s} {units[i]}')
    print('\nstalls (warps per issue slot):')
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
            stalls.append((float(row[i] or 0), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
    for v, h in sorted(stalls, reverse=True)[:10]:
        print(f'  {h:24s} {v:6.2f}')

    src = ncu_csv(report, 'source')
    # the source page repeats a two-line header per kernel; take the block of the requested launch
    starts = [i for i, r in enumerate(src) if r and r[0] == 'Kernel Name']
    per = max(1, len(starts) // max(1, len(raw) - 2))     # some ncu versions print every kernel's block twice
    k = launch * per
    begin = starts[k] if k < len(starts) else starts[0]
    end = starts[k + 1] if k + 1 < len(starts) else len(src)
    shdr, data = src[begin + 1], src[begin + 2:end]
    c_src, c_inst, c_smp = shdr.index('Source'), shdr.index('Instructions Executed'), shdr.index('# Samples')
    total = sum(float(r[c_inst] or 0) for r in data)
    samples = sum(float(r[c_smp] or 0) for r in data) or 1.0
    warps = float(data[0][c_inst] or 1)          # the first instruction runs once per warp
    print(f'\n{len(data)} SASS instructions, {total:.0f} executed, {samples:.0f} stall samples; '
          f'{warps:.0f} warps launched')

    def opcode(text):
        parts = text.split()
        if not parts:
            return '?'
        return (parts[1] if parts[0].startswith('@') and len(parts) > 1 else parts[0]).split('.')[0]

    runs, cur = [], None
    for idx, r in enumerate(data):
        count = float(r[c_inst] or 0)
        key = count / warps
        if cur and abs(cur['key'] - key) <= 0.15 * max(cur['key'], 1.0):
            cur['n'] += 1
            cur['inst'] += count
            cur['smp'] += float(r[c_smp] or 0)
            cur['ops'][opcode(r[c_src])] += 1
        else:
            cur = {'start': idx, 'key': key, 'n': 1, 'inst': count, 'smp': float(r[c_smp] or 0),
                   'ops': collections.Counter([opcode(r[c_src])])}
            runs.append(cur)
    print('\nSASS ranges by execution count per warp (>= 1.5 % of the executed instructions):')
    for x in runs:
        if x['inst'] / total >= 0.015:
            mix = ', '.join(f'{k} {v}' for k, v in x['ops'].most_common(8))
            print(f"  @{x['start']:5d} +{x['n']:4d}  x{x['key']:8.1f}/warp  inst {100 * x['inst'] / total:5.1f} %  "
                  f"samples {100 * x['smp'] / samples:5.1f} %  {mix}")
    mix = collections.Counter()
    smp = collections.Counter()
    for r in data:
        mix[opcode(r[c_src])] += float(r[c_inst] or 0)
        smp[opcode(r[c_src])] += float(r[c_smp] or 0)
    print('\ninstruction mix:')
    for k, v in mix.most_common(16):
        print(f'  {k:12s} {100 * v / total:5.1f} %   samples {100 * smp[k] / samples:5.1f} %')


if __name__ == '__main__':
    main()
