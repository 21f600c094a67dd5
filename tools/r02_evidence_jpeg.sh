#!/bin/bash
# The JPEG-front-end part of tools/r02_evidence.sh alone (after a change to those kernels), plus the default bench line and
# the GPU test log.
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/r02_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/r02_bench.json 2> $O/r02_bench.err
python tools/jpeg_profile.py 512 > $O/r02_jpeg_profile.txt 2>&1
ncu --set full --clock-control none -k regex:"jhuff|jpeg_idct|jpeg_color" -s 16 -c 8 -o /tmp/r02_full_jpeg -f python tools/jpeg_profile.py 512 > $O/r02_ncu_full3.log 2>&1
{
  echo "# ncu --set full --clock-control none, the launches of one stage of tools/jpeg_profile.py 512 (JPEG in, Huffman decoding on the GPU)"
  python tools/ncu_brief.py /tmp/r02_full_jpeg.ncu-rep
} > $O/r02_ncu_jpeg_per_launch.txt 2>&1
ncu --set full --clock-control none -k regex:"jenc|jpeg_enc|jpeg_fdct|draw_overlay" -s 14 -c 7 -o /tmp/r02_full_enc -f python tools/worker_profile.py 64 > $O/r02_ncu_full4.log 2>&1
{
  echo "# ncu --set full --clock-control none, the encoder-side launches of one 64-frame chunk of tools/worker_profile.py (uf_worker_batch_jpeg: MJPG in, annotated JPEG out)"
  python tools/ncu_brief.py /tmp/r02_full_enc.ncu-rep
} > $O/r02_ncu_jpeg_enc_per_launch.txt 2>&1
tail -3 $O/r02_pytest_gpu.log
