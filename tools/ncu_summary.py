"""Print the key metrics and the top stalled SASS lines of an .ncu-rep (run here, no GPU needed)."""
import csv, subprocess, sys, io
KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum',
        'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'sm__cycles_active.avg']
def main(path, top=22):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    r = rows[2]
    print('kernel:', r[hdr.index('Kernel Name')][:90])
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k); print(f'  {k} = {r[i]} {units[i]}')
    for i, h in enumerate(hdr):
        if 'tensor' in h and 'pct' in h and r[i] not in ('0', ''):
            print(f'  {h} = {r[i]} {units[i]}')
    raw = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r)
    hdr = rows[hi]
    ia, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    data = []
    for idx, r in enumerate(rows[hi + 1:]):
        if len(r) > isamp and r[isamp].isdigit():
            st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
            data.append((int(r[isamp]), idx, r[ia].strip()[:70], int(r[iex] or 0), st))
    tot = sum(d[0] for d in data)
    print(f'  total samples {tot}')
    for s, idx, src, ex, st in sorted(data, reverse=True)[:top]:
        print(f'  {100*s/tot:5.1f}% #{idx:4d} ex={ex:8d} {src:70s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}')
if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 22)
