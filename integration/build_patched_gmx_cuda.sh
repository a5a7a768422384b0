#!/bin/bash
# Builds the reference (fork) WITH its CUDA back end for sm_100 and the fepb200 hooks: the fork's own `mdrun -nb gpu -fep gpu`,
# `mdrun -nb gpu -fep cpu` + GMX_FEPB200 (our library beside its GPU non-bonded kernels, host hand-over) and
# `mdrun -nb gpu -fep gpu` + GMX_FEPB200 (our library inside its GPU route, device-resident) from one binary.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC=/tmp/gmxsrc
BUILD=/tmp/gmxbuild_cuda
OUT="$ROOT/integration/_gmx_cuda"
if [ ! -d "$SRC/src" ]; then
  rm -rf "$SRC"; mkdir -p "$SRC"; cp -r /root/reference/. "$SRC/"
fi
# (re-)apply the hook to a pristine copy of the one file it touches
cp /root/reference/src/gromacs/nbnxm/freeenergydispatch.cpp "$SRC/src/gromacs/nbnxm/freeenergydispatch.cpp"
(cd "$SRC" && patch -p1 < "$ROOT/integration/gromacs_shim/freeenergydispatch_fepb200.patch")
# ... and the hook for the perturbed 1-4 pairs (listed_forces/pairs.cpp -> fepb200_pairs14_*)
cp /root/reference/src/gromacs/listed_forces/pairs.cpp "$SRC/src/gromacs/listed_forces/pairs.cpp"
(cd "$SRC" && patch -p1 < "$ROOT/integration/gromacs_shim/pairs_fepb200.patch")
# ... and the scope around the foreign-lambda loop of the listed forces (all points of the 1-4 pairs in one library call)
cp /root/reference/src/gromacs/listed_forces/listed_forces.cpp "$SRC/src/gromacs/listed_forces/listed_forces.cpp"
(cd "$SRC" && patch -p1 < "$ROOT/integration/gromacs_shim/listed_forces_fepb200.patch")
# ... and the hooks in the fork's GPU route (mdrun -nb gpu -fep gpu + GMX_FEPB200: libfepb200 instead of k_calc_nb_fep*)
for f in src/gromacs/nbnxm/nbnxm_gpu_data_mgmt.cpp src/gromacs/nbnxm/cuda/nbnxm_cuda.cu src/gromacs/mdlib/sim_util.cpp; do cp "/root/reference/$f" "$SRC/$f"; done
(cd "$SRC" && patch -p1 < "$ROOT/integration/gromacs_shim/nbnxm_gpu_fepb200.patch")
# ... and, on top of them, the hooks for the cluster-pair kernel of the library (GMX_FEPB200_NB: fepb200_nb_* instead of nbnxn_kernel_*_cuda)
(cd "$SRC" && patch -p1 < "$ROOT/integration/gromacs_shim/nbnxm_gpu_nb_fepb200.patch")
mkdir -p "$BUILD"
cmake -G Ninja -S "$SRC" -B "$BUILD" -DCMAKE_C_COMPILER=/usr/bin/gcc -DCMAKE_CXX_COMPILER=/usr/bin/g++ \
  -DCMAKE_POLICY_VERSION_MINIMUM=3.5 -DGMX_GPU=CUDA -DGMX_CUDA_TARGET_SM=100 -DCUDA_TOOLKIT_ROOT_DIR=/usr/local/cuda -DCMAKE_CUDA_COMPILER=/usr/local/cuda/bin/nvcc -DGMX_MPI=OFF -DGMX_THREAD_MPI=ON -DGMX_OPENMP=ON \
  -DGMX_FFT_LIBRARY=fftpack -DGMX_DOUBLE=OFF -DGMX_HWLOC=OFF -DGMX_EXTERNAL_BLAS=OFF -DGMX_EXTERNAL_LAPACK=OFF \
  -DGMXAPI=OFF -DBUILD_TESTING=OFF -DCMAKE_BUILD_TYPE=Release -DGMX_SIMD=AVX2_256 -DGMX_CYCLE_SUBCOUNTERS=ON \
  "-DCMAKE_CXX_FLAGS=-I$ROOT/include -I$ROOT/integration/gromacs_shim" > "$BUILD/cmake.log" 2>&1
ninja -C "$BUILD" gmx > "$BUILD/ninja.log" 2>&1
mkdir -p "$OUT/bin" "$OUT/lib"
cp "$BUILD/bin/gmx" "$OUT/bin/"
cp -P "$BUILD"/lib/libgromacs.so* "$OUT/lib/"
cp -P "$BUILD"/lib/libmuparser.so* "$OUT/lib/" 2>/dev/null || true
strip "$OUT/lib/"*.so.*.* 2>/dev/null || true
# run inputs of the reference's free-energy test systems
TPR="$ROOT/tests/golden/mdrun_tpr"; mkdir -p "$TPR"
export GMXLIB="$SRC/share/top" LD_LIBRARY_PATH="$OUT/lib"
for sys in coulandvdwsequential_coul coulandvdwsequential_vdw coulandvdwtogether transformAtoB vdwalone; do
  d="$SRC/src/testutils/simulationdatabase/freeenergy/$sys"
  [ -f "$TPR/$sys.tpr" ] && continue   # committed run inputs are kept (they are what the GPU runs were made with)
  (cd /tmp && "$OUT/bin/gmx" -quiet grompp -f "$d/grompp.mdp" -c "$d/conf.gro" -p "$d/topol.top" -o "$TPR/$sys.tpr" \
     -po /tmp/mdout_$sys.mdp -maxwarn 10 > /tmp/grompp_$sys.log 2>&1) || { echo "grompp failed for $sys"; tail -5 /tmp/grompp_$sys.log; }
done
ls -la "$OUT/bin" "$OUT/lib" "$TPR"
