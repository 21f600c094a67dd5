#!/bin/bash
# compute-sanitizer over tools/sanitize_target.py on the GPU box; the summary lands in gpurun_out/r02_sanitizer_summary.txt
O=gpurun_out/r02_sanitizer_summary.txt
{
  echo "# compute-sanitizer on tools/sanitize_target.py (every round-2 kernel once: fused resize+stem incl. border tiles, NMS bit matrix + sweep"
  echo "# with 700 / 4420 / 5001 candidates and a whole batch of all-prior frames, JPEG decode with Huffman decoding on the device and on the host —"
  echo "# 4:4:4 / 4:2:2 / 4:2:0 / grey / odd sizes / truncated / bit-flipped / multi-CTA frames —, rectangles + text overlay + JPEG encode incl."
  echo "# dummy-block geometries, per frame (host Huffman coder) and in the batch forms (Huffman coder + byte stuffing on the device, the"
  echo "# worker call), the batcher with RGB and JPEG frames), final round-2 build"
  echo "compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_target.py"
  compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_target.py 2>&1 | grep -E "sanitize target|ERROR SUMMARY|Invalid|Error|error" | head -20
  echo "memcheck rc=${PIPESTATUS[0]}"
  echo "compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_target.py"
  compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_target.py 2>&1 | grep -E "sanitize target|RACECHECK SUMMARY|hazard|Error|error" | head -20
  echo "racecheck rc=${PIPESTATUS[0]}"
} > $O 2>&1
cat $O
