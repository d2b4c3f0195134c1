"""Condense .ncu-rep files (on the GPU box, where they were written) into small text summaries that fit gpurun's 64 MiB
return limit:   python profiles/tools/ncu_summarize.py OUT.txt REP [REP ...]
Per report: the raw metrics the roofline / bound statements in DESIGN.md cite, the stall-reason totals of the warp
samples, and the hottest SASS lines (instructions executed, samples, dominant stall)."""
import csv, io, subprocess, sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
           "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum"]


def run(args):
    return subprocess.run(["ncu", "-i", *args], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main(out, reps):
    with open(out, "w") as fo:
        for rep in reps:
            rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
            if len(rows) < 3:
                fo.write(f"== {rep}: no data\n")
                continue
            hdr, units = rows[0], rows[1]
            ix = {n: i for i, n in enumerate(hdr)}
            for r in rows[2:]:
                fo.write(f"== {r[ix['Kernel Name']][:90]} grid {r[ix.get('Grid Size', 0)]} block {r[ix.get('Block Size', 0)]}  [{rep.split('/')[-1]}]\n")
                for m in METRICS:
                    if m in ix:
                        fo.write(f"   {m:<84} {units[ix[m]]:<12} {r[ix[m]]}\n")
            src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
            # first block = SASS view of the first kernel
            start = next((i for i, r in enumerate(src) if r and r[0] == "Address"), None)
            if start is None:
                continue
            h = src[start]
            ix = {n: i for i, n in enumerate(h)}
            stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
            data = []
            for r in src[start + 1:]:
                if not r or r[0] in ("Kernel Name", "Address") or len(r) < len(h):
                    break
                try:
                    data.append((r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]]),
                                 {s[6:]: int(r[ix[s]] or 0) for s in stalls}))
                except ValueError:
                    continue
            ti, ts = sum(d[1] for d in data) or 1, sum(d[2] for d in data) or 1
            tot = {}
            for d in data:
                for k, v in d[3].items():
                    tot[k] = tot.get(k, 0) + v
            sv = sum(tot.values()) or 1
            fo.write(f"   -- SASS: {len(data)} instructions, {ti} executed (warp level), {ts} samples; stall reasons: "
                     + ", ".join(f"{k} {100 * v / sv:.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v > 0.02 * sv) + "\n")
            for i in sorted(sorted(range(len(data)), key=lambda i: -data[i][2])[:14]):
                d = data[i]
                top = max(d[3].items(), key=lambda kv: kv[1]) if d[3] else ("", 0)
                fo.write(f"   hot {i:5d}  {d[0][:64]:<64} exec {d[1]:>9}  samples {100 * d[2] / ts:5.1f}%  ({top[0]})\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
