#!/bin/bash
mkdir -p gpurun_out
export JD_LIB_PATH=$PWD/jolideco_b200/libjolideco_b200_trace.so
timeout 200 python tools/tcm_trace.py 1024 5 2>&1 | tail -34
