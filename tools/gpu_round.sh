#!/bin/bash
# One GPU-box call: parity tests, the bench line, its ncu launch list, sweeps, ncu --set full captures.
# Usage (from the repo root on the GPU box):  bash tools/gpu_round.sh [tag]
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_reference_arm.json 2>> $out/${tag}_bench_n1.err
timeout 600 python tools/in_sweep.py --out $out/${tag}_in_sweep.txt > /dev/null 2> $out/${tag}_in_sweep.err
timeout 900 python tools/sweep.py programs nulls --rows 600037902 --out $out/${tag}_sweep_programs_600M_rows.jsonl > /dev/null 2> $out/${tag}_sweep_programs.err
timeout 300 python tools/narrow_unpack_probe.py > $out/${tag}_narrow_unpack.txt 2>&1
timeout 300 python tools/codec_cold_probe.py > $out/${tag}_codec_cold.txt 2>&1
timeout 300 python tools/dict_probe.py > $out/${tag}_dict_probe.txt 2>&1
timeout 600 python tools/sweep.py aux --rows 268435456 --out $out/${tag}_sweep_aux.jsonl > /dev/null 2> $out/${tag}_sweep_aux.err
# launch list of the bench command (shares, not absolutes)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_bench_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu > $out/${tag}_bench_under_ncu.log 2>&1
( cd impala-avx2-parquet-scanner_b200/host && timeout 600 ./tests/scanner-bench --json --threads 1,2,3,4 ) > $out/${tag}_scanner_bench.json 2> $out/${tag}_scanner_bench.err
for c in between12 q6agg in256 in256x8 q6nulls; do
  sel="-c 2"
  # (q6nulls: the first call also runs the optimistic no-NULL scan -- 4 matching launches; the second one is tile counts, prefixes, chain)
  [ $c = q6nulls ] && sel="-s 4 -c 3"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'eval|fle_jit|null_chain|nc_prefix|def_tile' $sel -f -o $out/${tag}_${c} \
      python tools/profile_one.py $c 600037902 3 > $out/${tag}_ncu_${c}.log 2>&1
  # only the summary travels back (gpurun_out/ is capped at 64 MiB; a report with sources is 10-40 MB)
  python tools/ncu_brief.py $out/${tag}_${c}.ncu-rep > $out/${tag}_${c}_ncu_full_summary.txt 2>&1
  rm -f $out/${tag}_${c}.ncu-rep
done
ls -la $out
