#!/bin/bash
# tools/check_sass.sh [lib.so]: the kernels that carry UpdateMatrices (k_um0, k_iter, k_iter64, k_update_matrices) promise cv2's
# UNCONTRACTED float arithmetic.  Their packed f32x2 multiplies are safe, but ptxas contracts a packed multiply feeding a packed
# add into FFMA2 even under .rn / -fmad=false (um_device.cuh), so no FFMA2 -- and no scalar FFMA outside the reciprocal's Newton
# step (__frcp_rn: 2 per call site) and cartToPolar-style explicit fmaf() -- may appear in them.  Prints the offenders; exit 1 if any.
lib=${1:-$(dirname "$0")/../optical_flow_b200/lib/libofb200.so}
command -v cuobjdump > /dev/null || { echo "cuobjdump not found"; exit 2; }
cuobjdump -sass "$lib" 2>/dev/null | awk '
  /Function :/ { f = $3 }
  /FFMA2/ && f ~ /k_um0|k_iterI|k_iter64|k_update_matrices/ { c[f]++ }
  END { n = 0; for (k in c) { print c[k], "FFMA2 in", k; n++ } exit n ? 1 : 0 }'
