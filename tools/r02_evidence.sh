#!/bin/bash
# Round-2 evidence collection on the GPU box (run through gpurun). Everything lands in gpurun_out/ as SMALL text files (the
# .ncu-rep captures are summarised on the box and deleted: gpurun copies back at most 64 MiB); tools/r02_collect.py then
# copies them into profiles/.
set -x
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/r02_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/r02_bench.json 2> $O/r02_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err
python bench.py --steps 10 --warmup 3 --cls-bias 3 --no-cpu-baseline --latency-iters 20 > $O/r02_bench_cfg5_rfb320_allpriors.json 2>> $O/r02_bench.err
for pb in "p1 -1.5" "p5 -0.75" "p50 0.65"; do set -- $pb
  python bench.py --steps 10 --warmup 3 --net 640x480 --batch 64 --cls-bias $2 --no-cpu-baseline --latency-iters 20 > $O/r02_bench_cfg5_rfb640_b64_$1.json 2>> $O/r02_bench.err
done
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs --latency-iters 5 > $O/r02_ncu_launches.log 2>&1
# full captures of the kernels of one stage (graphs off), summarised here
ncu --set full --clock-control none -s 40 -c 44 -o /tmp/r02_full_stage -f python tools/profile_target.py 2 320x240 256 2 > $O/r02_ncu_full1.log 2>&1
ncu --set full --clock-control none -k regex:"nms_mask|nms_sweep|post_kernel" -s 3 -c 3 -o /tmp/r02_full_nms -f python tools/profile_target.py 2 320x240 256 2 3.0 > $O/r02_ncu_full2.log 2>&1
{
  echo "# ncu --set full --clock-control none, one line per captured launch (tools/ncu_brief.py), round 2, final build"
  echo "# r02_full_stage: the launches of one 128-frame stage of the benchmark (tools/profile_target.py, CUDA graphs off)"
  echo "# r02_full_nms: post / bit-matrix / sweep kernels with every prior a candidate (cls_bias 3)"
  echo "## r02_full_stage"; python tools/ncu_brief.py /tmp/r02_full_stage.ncu-rep
  echo "## r02_full_nms"; python tools/ncu_brief.py /tmp/r02_full_nms.ncu-rep
} > $O/r02_ncu_full_per_launch.txt 2>&1
python tools/ncu_traffic.py /tmp/r02_full_stage.ncu-rep /tmp/r02_full_nms.ncu-rep > $O/r02_ncu_dram_bytes.csv 2>> $O/r02_evidence_err.log
# the JPEG front end (N2): where its time goes (event pairs per kernel family), and full captures of one 128-frame stage's
# Huffman / IDCT / colour launches
python tools/jpeg_profile.py 512 > $O/r02_jpeg_profile.txt 2>&1
ncu --set full --clock-control none -k regex:"jhuff|jpeg_idct|jpeg_color" -s 16 -c 8 -o /tmp/r02_full_jpeg -f python tools/jpeg_profile.py 512 > $O/r02_ncu_full3.log 2>&1
{
  echo "# ncu --set full --clock-control none, the launches of one stage of tools/jpeg_profile.py 512 (JPEG in, Huffman decoding on the GPU)"
  python tools/ncu_brief.py /tmp/r02_full_jpeg.ncu-rep
} > $O/r02_ncu_jpeg_per_launch.txt 2>&1
ncu -i /tmp/r02_full_stage.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]; ik=hdr.index('Kernel Name')
cols=[i for i,h in enumerate(hdr) if 'pipe_tensor' in h][:10]
print('# tensor-pipe metrics of the tcgen05 kernels (ncu --set full, one stage of the benchmark)')
print('kernel | '+' | '.join(hdr[i]+' ['+rows[1][i]+']' for i in cols))
for r in rows[2:]:
    if 'tc_kernel' in r[ik]: print(r[ik][:52]+' | '+' | '.join(r[i] for i in cols))
" > $O/r02_ncu_tensor_pipe.txt 2>> $O/r02_evidence_err.log
ls -la $O | tail -30
