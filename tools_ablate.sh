#!/bin/bash
# usage: tools_ablate.sh lib1.so lib2.so ...  — ms per launch of the PCG mat-vec (time_matvec hook) for each build, on one settled 1M state
mkdir -p gpurun_out
export VFD_ABLATE_STATE=/tmp/vfd_ablate_state.npz
python tools_ablate.py
for lib in "$@"; do VFD_LIB=$lib python tools_ablate.py; done
