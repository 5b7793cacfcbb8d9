#!/bin/bash
mkdir -p gpurun_out
timeout 300 tools/probe_exchange 400 > gpurun_out/r2s5_exchange.log 2>&1; grep -c mismatch gpurun_out/r2s5_exchange.log; grep mismatch gpurun_out/r2s5_exchange.log | head -40
