#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/umma_rate.py > gpurun_out/umma_rate.log 2>&1; echo "rc=$?"; cat gpurun_out/umma_rate.log
