#!/bin/bash
# round 2, session 4: flags-in-data decode hand-off vs the bulk-copy queue depth (probe_exchange says inflight 3 costs 8-22 us per exchange)
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" GVL_MEGA_TRACE=1 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/r2s4_$name.log 2>&1
  head -1 gpurun_out/r2s4_$name.log
}
run ll0_cur3 GVL_MEGA_LL=0
run ll0_cur1 GVL_MEGA_LL=0 GVL_MEGA_INFLIGHT_CUR=1
run ll1_cur3 GVL_MEGA_LL=1
run ll1_cur1 GVL_MEGA_LL=1 GVL_MEGA_INFLIGHT_CUR=1
run ll1_cur2 GVL_MEGA_LL=1 GVL_MEGA_INFLIGHT_CUR=2
run ll1_if2_cur2 GVL_MEGA_LL=1 GVL_MEGA_INFLIGHT=2 GVL_MEGA_INFLIGHT_CUR=2
