"""Summarise an .ncu-rep: key metrics + hot regions of the SASS by executed instructions and stall samples.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--regions]"""
import csv, subprocess, sys, io

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread ',
        'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor', 'launch__occupancy_limit', 'smsp__inst_executed.sum ',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg', 'sm__pipe_alu_cycles_active.avg', 'sm__pipe_fmaheavy_cycles_active.avg', 'sm__pipe_fmalite_cycles_active.avg', 'sm__inst_executed_pipe_fmaheavy', 'sm__inst_executed_pipe_fmalite', 'sm__inst_executed_pipe_lsu.avg.pct',
        'sm__inst_executed_pipe_xu.avg.pct', 'sm__inst_executed_pipe_uniform.avg.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__average_warps_issue_stalled', 'sm__cycles_elapsed.max ', 'lts__t_bytes.sum ',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ', 'launch__shared_mem_per_block_dynamic', 'dram__throughput', 'lts__t_sectors_srcunit_tex_op_read.sum ', 'lts__t_sectors_srcunit_tex_op_write.sum ']


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print('== kernel:', vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
        for h, u, v in zip(hdr, units, vals):
            if any((h + ' ').startswith(k) or k.strip() in h and k.endswith('stalled') or (k.strip() in h and not k.endswith(' ')) for k in KEYS):
                if 'issue_stalled' in h and 'per_issue_active' not in h:
                    continue
                print(f'  {h:95s} {u:12s} {v}')


def regions(path, thresh=0.01):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # may contain several kernels: split on "Kernel Name" lines
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}; blocks.append(cur)
        elif cur is not None:
            cur['rows'].append(r)
    for b in blocks:
        hdr, data = b['rows'][0], b['rows'][1:]
        isrc, iex, ist = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
        tot = sum(int(r[iex]) for r in data); tst = sum(int(r[ist]) for r in data) or 1
        print('== source regions:', b['name'][:100], 'instructions', len(data), 'executed', tot, 'samples', tst)
        regs = []
        for i, r in enumerate(data):
            ex, st = int(r[iex]), int(r[ist])
            if regs and regs[-1][2] == ex:
                regs[-1][1] = i; regs[-1][3] += st; regs[-1][4] += ex
            else:
                regs.append([i, i, ex, st, ex])
        for a, bb, ex, st, sumex in regs:
            if sumex > tot * thresh or st > tst * thresh:
                print(f'  [{a:4d}-{bb:4d}] n={bb-a+1:4d} exec/instr={ex:9d} share={100*sumex/tot:5.1f}% stall={100*st/tst:5.1f}%  {data[a][isrc].strip()[:70]}')


if __name__ == '__main__':
    raw(sys.argv[1])
    if '--regions' in sys.argv:
        regions(sys.argv[1])
