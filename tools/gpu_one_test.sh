#!/bin/bash
# Run the tests selected by $K (pytest -k expression) on the GPU box.
mkdir -p gpurun_out
timeout ${TMO:-600} python -m pytest tests -m gpu -q --timeout 300 -k "$K" > gpurun_out/one_test.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-40} gpurun_out/one_test.log | cut -c1-300
