#!/bin/bash
bash tools/gpu_multi.sh 2
bash tools/gpu_bench.sh 2
