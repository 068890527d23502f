"""Summarise an .ncu-rep: headline metrics per launch + top stall sites (needs -lineinfo + --import-source on)."""
import csv, io, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.avg', 'lts__t_sector_hit_rate.pct', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_uniform.sum',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']

def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout

def main(path, top=22, dump=None):
    rows = list(csv.reader(io.StringIO(run([path, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index('Kernel Name')][-90:])
        for i, h in enumerate(hdr):
            if h in KEYS:
                print(f"   {h:75s} {r[i]:>16s} {units[i]}")
    rows = list(csv.reader(io.StringIO(run([path, "--page", "source", "--csv"]))))
    # the source page repeats a 2-line header per kernel
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == 'Kernel Name':
            name = rows[i][1]; hdr = rows[i + 1]; j = i + 2
            data = []
            while j < len(rows) and not (rows[j] and rows[j][0] == 'Kernel Name'):
                data.append(rows[j]); j += 1
            si, src = hdr.index('# Samples'), hdr.index('Source')
            if dump:
                ie, wf = hdr.index('Instructions Executed'), hdr.index('L1 Wavefronts Shared')
                with open(dump, 'a') as fh:
                    fh.write(f"## {name}\n#idx\tinst_exec\tsamples\twavefronts_shared\tsass\tstalls\n")
                    stall_cols = [c for c, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
                    for k, r in enumerate(data):
                        st = sorted(((hdr[c][6:], int(r[c])) for c in stall_cols if r[c] not in ('', '0')), key=lambda kv: -kv[1])[:3]
                        fh.write(f"{k}\t{r[ie] or 0}\t{r[si] or 0}\t{r[wf] or 0}\t{r[src].strip()[:70]}\t{st}\n")
            stall = [c for c, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
            tot = sum(int(r[si] or 0) for r in data)
            print(f"-- stall sites of {name[-70:]} (total samples {tot})")
            for k in sorted(sorted(range(len(data)), key=lambda k: -int(data[k][si] or 0))[:top]):
                r = data[k]
                st = sorted(((hdr[c][6:], int(r[c])) for c in stall if r[c] not in ('', '0')), key=lambda kv: -kv[1])[:2]
                print(f"   {k:5d} {int(r[si]):6d} {100*int(r[si])/max(tot,1):5.1f}%  {r[src].strip()[:78]:78s} {st}")
            i = j
        else:
            i += 1

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 22, sys.argv[3] if len(sys.argv) > 3 else None)
