#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python tools/time_inference.py --config ljspeech --text-len 120 --frames 640 > $OUT/time_inference_ljspeech.log 2>&1; echo "rc=$?"; tail -4 $OUT/time_inference_ljspeech.log
timeout 600 python tools/time_inference.py --config generated_training --text-len 120 --frames 640 > $OUT/time_inference_generated.log 2>&1; echo "rc=$?"; tail -4 $OUT/time_inference_generated.log
timeout 600 python tools/time_inference.py --config ljspeech --text-len 120 --frames 640 --precision fp32 > $OUT/time_inference_ljspeech_fp32.log 2>&1; echo "rc=$?"; tail -2 $OUT/time_inference_ljspeech_fp32.log
