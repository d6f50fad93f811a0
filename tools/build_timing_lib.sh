#!/bin/bash
# Measurement build of the team / duo kernels (phase counters, DEB_TICK): links a second library next to the product's objects.
set -e
cd "$(dirname "$0")/../disco-eb_b200/csrc"
make -s
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DDEB_TEAM_TIMING $EXTRA -c -o _obj/deb_team_timing.o deb_team.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o _obj/libdeb_timing.so _obj/deb_kernels.o _obj/deb_team_timing.o _obj/deb_lane.o _obj/deb_background.o _obj/deb_dist.o _obj/deb_spectra.o -ldl
ls -la _obj/libdeb_timing.so
