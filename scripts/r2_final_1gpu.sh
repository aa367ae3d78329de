# Round-2 evidence on one B200: GPU tests, every bench workload, the reference arm, launch list, step timeline.
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2final; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt; tail -3 $O/pytest_gpu.txt
timeout 600 python bench.py > $O/bench_bf16.json 2> $O/bench_bf16.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; echo "ref rc=$?"
timeout 600 python bench.py --host-dtype fp32 --no-cpu-baseline --no-strong > $O/bench_bf16_fp32host.json 2> /dev/null; echo "fp32host rc=$?"
timeout 600 python bench.py --shape scaled > $O/bench_scaled.json 2> $O/bench_scaled.err; echo "scaled rc=$?"
timeout 600 python bench.py --workload infer_sweep > $O/bench_infer_sweep.json 2> $O/bench_infer_sweep.err; echo "sweep rc=$?"
timeout 600 python bench.py --workload infer_sweep --no-mask-hint > $O/bench_infer_sweep_dense.json 2> /dev/null; echo "sweep dense rc=$?"
timeout 600 python bench.py --workload ece > $O/bench_ece.json 2> $O/bench_ece.err; echo "ece rc=$?"
timeout 900 python bench.py --workload raw_infer > $O/bench_raw_infer.json 2> $O/bench_raw_infer.err; echo "raw rc=$?"
TL=$PWD/multimodal-sensor-fusion-with-attention-rajeevatla_b200/libmsf_b200_timeline.so
MSF_B200_LIB=$TL timeout 300 python scripts/step_timeline.py 4096 > $O/step_timeline.txt 2>&1; echo "timeline rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong --steps-per-graph 1 > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong --steps-per-graph 1 > $O/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python scripts/wg_prof_target.py > $O/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"chain2_kernel|head_kernel|proj_kernel|wg2_kernel|opt_pack_kernel" -s 12 -c 6 -o $O/top_kernels -f python scripts/wg_prof_target.py > $O/ncu_full.log 2>&1
echo "ncu full rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2final/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d.get("e2e",{}); r=d.get("roofline",{})
        print(f.split('/')[-1], "ms", round(d.get("ms_per_step",0),4), "value", round(d.get("value",0)), d.get("unit"), "e2e", round(e.get("value",0)), "frac", r.get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex: print(f, "ERR", ex)
PY
