"""Turn the raw ncu artefacts a gpurun visit brought back (gpurun_out/) into the small tracked summaries under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches_tf32.csv profiles/r1_launches_tf32.md "title"
    python tools/summarize_profiles.py report   gpurun_out/prof_grid.ncu-rep  profiles/r1_ncu_grid_sample.md "title"
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'launch__waves_per_multiprocessor',
        'sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio']


def launches(src, dst, title):
    rows = list(csv.reader(open(src)))
    for i, r in enumerate(rows):
        if 'Kernel Name' in r:
            hdr, start = r, i + 1
            break
    ki, mi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for r in rows[start:]:
        if len(r) <= mi:
            continue
        name = re.sub(r'\(.*', '', r[ki]).replace('void ', '').replace('<unnamed>::', '')
        v = float(r[mi].replace(',', '')) / (1000.0 if r[ui] == 'ns' else 1.0)
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    n = sum(a[0] for a in agg.values())
    with open(dst, 'w') as f:
        f.write('# %s\n\nSource: `ncu --metrics gpu__time_duration.sum --clock-control none` launch list (%s).\n'
                'Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n'
                'launches: %d, summed device time: %.1f us\n\n| kernel | launches | us | share |\n|---|---:|---:|---:|\n'
                % (title, src, n, tot))
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write('| `%s` | %d | %.1f | %.1f %% |\n' % (k[:90], c, t, 100 * t / tot))


def report(src, dst, title):
    out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, 'w') as f:
        f.write('# %s\n\nSource: `ncu --set full --clock-control none --import-source on` (%s), read with '
                '`ncu -i ... --page raw --csv`.\n\n' % (title, src))
        for r in rows[2:]:
            f.write('## %s  grid %s\n\n| metric | value | unit |\n|---|---:|---|\n'
                    % (re.sub(r'\(.*', '', r[hdr.index('Kernel Name')]), r[hdr.index('Grid Size')] if 'Grid Size' in hdr else ''))
            for k in KEYS:
                if k in hdr:
                    f.write('| %s | %s | %s |\n' % (k, r[hdr.index(k)], units[hdr.index(k)]))
            f.write('\n')


if __name__ == '__main__':
    {'launches': launches, 'report': report}[sys.argv[1]](*sys.argv[2:5])
