#!/usr/bin/env bash
# Builds the CUDA backend (libgknext_cuda.so, sm_100a only) and the host mirror
# (libgknext_host.so) in-tree.  nvcc cross-compiles without a GPU.
set -euo pipefail
cd "$(dirname "$0")"
mkdir -p lib build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
# -fmad=false: multiply-adds are only fused where the source writes fmaf() (see gk_common.cuh)
NVFLAGS="-O3 -std=c++17 -lineinfo -fmad=false $ARCH -Xcompiler -fPIC -Xcompiler -fno-strict-aliasing ${GK_NVCC_EXTRA:-}"
objs=()
pids=()
for f in gk_api gk_scene gk_bvh_build gk_integrator gk_filters gk_exchange gk_probes; do
  src=csrc/$f.cu; obj=build/$f.o
  objs+=("$obj")
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ -n "$(find csrc ../include -newer "$obj" \( -name '*.cuh' -o -name '*.h' \) -print -quit)" ]; then
    $NVCC $NVFLAGS -c "$src" -o "$obj" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
$NVCC $ARCH -shared -o lib/libgknext_cuda.so "${objs[@]}" -lcudart
g++ -O2 -std=c++17 -fPIC -shared -Wall -o lib/libgknext_host.so host/gk_assets.cpp host/gk_engine.cpp host/gk_host_capi.cpp \
    -pthread -Llib -lgknext_cuda -Wl,-rpath,'$ORIGIN'
# the multi-GPU compositor over NCCL (include/gknext_compositor.h); NCCL headers/libs: the system package of this image
COMP="g++ -O2 -std=c++17 -fPIC -shared -Wall -o lib/libgknext_comp.so host/gk_compositor.cpp -I/usr/local/cuda/include -Llib -lgknext_cuda -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,\$ORIGIN"
if ! $COMP -lnccl 2>/dev/null; then
  # no system NCCL dev package: link against the one PyTorch ships (same SONAME, libnccl.so.2)
  NCCL_DIR=$(python -c "import nvidia.nccl, os; print(os.path.dirname(nvidia.nccl.__file__))" 2>/dev/null || true)
  if [ -z "$NCCL_DIR" ]; then NCCL_DIR=$(python -c "import os, site; print(os.path.join(site.getsitepackages()[0], 'nvidia', 'nccl'))"); fi
  $COMP -I"$NCCL_DIR/include" "$NCCL_DIR/lib/libnccl.so.2"
fi
echo "built lib/libgknext_cuda.so lib/libgknext_host.so lib/libgknext_comp.so"
