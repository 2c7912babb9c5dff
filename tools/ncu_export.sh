#!/bin/bash
# usage: tools/ncu_export.sh <case> <reps> <skip> <count> <tag>
# One `ncu --set full` capture of tools/ncu_cases.py <case>, exported ON THE BOX to small text files under gpurun_out/
# (the .ncu-rep itself stays in /tmp: five of them blew gpurun's 64 MiB return limit in round 2, call A).
c=$1; reps=$2; skip=$3; cnt=$4; tag=$5
rep=/tmp/ncu_$tag
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${NCU_K:-fa_fwd_kernel|fa_fwd_wide|fa_bwd|fa_combine}" -s $skip -c $cnt -f -o $rep \
  python tools/ncu_cases.py $c $reps > gpurun_out/${tag}_ncu.log 2>&1
ncu -i $rep.ncu-rep --page raw --csv > /tmp/${tag}_raw.csv 2>/dev/null
# keep the metrics the roofline / DESIGN tables quote
python - "$tag" <<'PY'
import csv, sys
tag = sys.argv[1]
rows = list(csv.reader(open(f"/tmp/{tag}_raw.csv")))
hdr = rows[0]
keep = [i for i, h in enumerate(hdr) if any(k in h for k in (
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor",
    "sm__inst_executed_pipe_tensor", "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct", "launch__registers_per_thread",
    "launch__grid_size", "sm__warps_active.avg.pct", "smsp__inst_executed_pipe_xu", "sm__inst_executed_pipe_xu", "sm__pipe_xu",
    "sm__pipe_fma", "sm__pipe_alu", "sm__pipe_fmaheavy", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum ",
    "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak", "gpu__dram_throughput", "sm__inst_executed.sum",
    "smsp__issue_active.avg.pct", "sm__mio", "l1tex__lsu_writeback", "lts__t_sectors_srcunit_tex_op_read.sum"))]
with open(f"gpurun_out/{tag}_metrics.csv", "w") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in keep if i < len(r)])
PY
ncu -i $rep.ncu-rep --page source --csv > /tmp/${tag}_src.csv 2>/dev/null
python tools/ncu_top.py /tmp/${tag}_src.csv 45 > gpurun_out/${tag}_stalls.txt 2>&1
tail -1 gpurun_out/${tag}_ncu.log
