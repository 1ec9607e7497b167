"""Turn ncu CSV exports (read here, no GPU needed) into the small summaries committed under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches.csv   > profiles/rNN_launches.md
    python tools/summarize_ncu.py raw      gpurun_out/kernel_raw.csv > profiles/rNN_kernel.md

`launches` reads the `--metrics gpu__time_duration.sum --csv --log-file` launch list; `raw` reads
`ncu -i X.ncu-rep --page raw --csv`.
"""
import collections
import csv
import sys

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed.sum.per_cycle_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_tc.sum', 'sm__inst_executed_pipe_uniform.sum',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'smsp__average_warp_latency_per_inst_issued.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
]


def launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    agg = collections.OrderedDict()
    total = 0.0
    for r in csv.DictReader(lines):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        t = float(r['Metric Value'].replace(',', ''))
        if r.get('Metric Unit') == 'us':
            t *= 1e3
        elif r.get('Metric Unit') == 'ms':
            t *= 1e6
        name = r['Kernel Name']
        name = name[:name.index('(')] if '(' in name else name
        key = (name, r['Grid Size'], r['Block Size'])
        agg.setdefault(key, []).append(t)
        total += t
    print('| kernel | grid | block | launches | total us | avg us | share |')
    print('|---|---|---|---|---|---|---|')
    for (name, g, b), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f'| `{name}` | {g} | {b} | {len(v)} | {sum(v) / 1e3:.1f} | {sum(v) / len(v) / 1e3:.2f} | {sum(v) / total * 100:.1f}% |')
    print(f'\ntotal device time of the listed launches: {total / 1e3:.1f} us (ncu per-launch times are cold-cache and '
          'serialised: compare shares, not absolutes)')


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"### `{r[col['Kernel Name']]}`  grid {r[col['Grid Size']]} block {r[col['Block Size']]}\n")
        print('| metric | value | unit |')
        print('|---|---|---|')
        for k in KEYS:
            if k in col:
                print(f'| {k} | {r[col[k]]} | {units[col[k]]} |')
        print()


if __name__ == '__main__':
    {'launches': launches, 'raw': raw}[sys.argv[1]](sys.argv[2])
