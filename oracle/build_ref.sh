#!/bin/sh
# Compiles the REFERENCE's own rotated-IoU/NMS and RoI-aware pooling CUDA kernels, from the sources where they lie under
# /root/reference (never copied), plus a small host shim reproducing iou3d_nms.cpp:79-126 over raw
# device pointers -> oracle/_ref/libref_iou3d.so.  GPU-only oracle; skipped when the reference tree
# is absent (e.g. on the GPU box, which uses the prebuilt file that travelled with the snapshot).
set -e
cd "$(dirname "$0")"
REF=${PCDET_REFERENCE:-/root/reference}
SRC="$REF/pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu"
if [ ! -f "$SRC" ]; then echo "reference tree absent; keeping prebuilt oracle/_ref"; exit 0; fi
mkdir -p _ref
nvcc -O2 -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a \
     -o _ref/libref_iou3d.so "$SRC" ref_iou3d_shim.cu
echo "built oracle/_ref/libref_iou3d.so"
SRC2="$REF/pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu"
if [ -f "$SRC2" ]; then
    nvcc -O2 -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a \
         -o _ref/libref_roiaware.so "$SRC2" ref_roiaware_shim.cu
    echo "built oracle/_ref/libref_roiaware.so"
fi
