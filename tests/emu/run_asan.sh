#!/bin/sh
# Optional: the SIMT-emulator build of the kernel sources under AddressSanitizer
# (compute-sanitizer is closed on the GPU pool).  Catches out-of-bounds shared-memory /
# "device" buffer accesses of the kernel logic on the CPU box.
set -e
cd "$(dirname "$0")/../.."
g++ -O1 -g -std=c++17 -DPCL_EMU -fPIC -shared -fsanitize=address -fno-omit-frame-pointer -x c++ \
    polarcode_and_ldpc_b200/csrc/pcl_api.cu -I tests/emu -I polarcode_and_ldpc_b200/csrc -o /tmp/libpcl_emu_asan.so
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
    PCL_EMU_LIB=/tmp/libpcl_emu_asan.so python tests/emu/asan_cases.py
