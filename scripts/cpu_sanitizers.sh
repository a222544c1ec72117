#!/bin/bash
# Sanitizer evidence that can be produced without compute-sanitizer (closed on this GPU pool): the kernels' source under
# ASan / UBSan / TSan through the CUDA execution-model emulation, the staging pool under TSan / ASan, the oracle and the
# host side of libb200expect under ASan / UBSan.  Log -> profiles/r2_cpu_sanitizers.log
cd "$(dirname "$0")/.."
{
  echo "# $(date -u +%FT%TZ)  g++ $(g++ -dumpversion), $(python --version 2>&1)"
  python -m pytest tests/test_sanitizers.py -v -rs 2>&1
} | tee profiles/r2_cpu_sanitizers.log | tail -25
