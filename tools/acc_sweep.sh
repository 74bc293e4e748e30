#!/bin/bash
# geometry sweep of welford_bulk_kernel (needs a CCST_DEV build); usage: tools/acc_sweep.sh NxCxHxW "ppc,slots,warps ..."
shape=$1; shift
for plan in "$@"; do
  echo "== $shape acc plan $plan"
  CCST_ACC_PLAN=$plan python tools/op_bench.py --shapes $shape 2>&1 | grep welford
done
