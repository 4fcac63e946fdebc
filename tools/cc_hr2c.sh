#!/bin/bash
# compile the real-input Hilbert kernel's translation unit alone and report registers / spills / opcode mix
cd /root/repo/fftconv_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -cudart shared -Xptxas=-v -c -o ../lib/obj/fftconv_cuda_hr2c.o fftconv_cuda_hr2c.cu 2>&1 | grep -E "spill|Used|error"
cuobjdump -sass ../lib/obj/fftconv_cuda_hr2c.o | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s+\/\*.*//' > /tmp/w/hr.lst
wc -l < /tmp/w/hr.lst
awk '{print $2}' /tmp/w/hr.lst | sed -E 's/\..*//' | sort | uniq -c | sort -rn | head -${1:-8} | tr '\n' ' '; echo
