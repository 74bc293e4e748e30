#!/bin/bash
# ring-geometry sweep of plane_bulk_kernel (needs a CCST_DEV build: python -m ccst_b200.build --force --dev)
# usage: tools/bulk_sweep.sh HxW "slots,warps,ppc ..."
shape=$1; shift
for plan in "$@"; do
  echo "== $shape plan $plan"
  CCST_BULK_PLAN=$plan python tools/op_bench.py --shapes $shape 2>&1 | grep -v "^op" | grep -v welford
done
