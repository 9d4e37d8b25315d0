#!/bin/bash
# development probe: transposed-operand int8 GEMM (no staging kernel) vs staged, and parity of the result
python tools/prof_rot.py 10000 25088 10 2>&1 | tail -1 | grep -o "'convert_ms': [0-9.]*\|'rotate_ms': [0-9.]*\|'reml_ms': [0-9.]*" | tr '\n' ' '; echo " (staged)"
PG_GEMM_TT=1 python tools/prof_rot.py 10000 25088 10 2>&1 | tail -1 | grep -o "'convert_ms': [0-9.]*\|'rotate_ms': [0-9.]*\|'reml_ms': [0-9.]*" | tr '\n' ' '; echo " (TT)"
PG_GEMM_TT=1 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "i8split or end_to_end or seeded" 2>&1 | tail -2
