#!/bin/bash
cd tools/ubench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fhfma_rate fhfma_rate.cu && timeout 120 ./fhfma_rate
