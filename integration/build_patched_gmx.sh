#!/bin/bash
# Builds the reference's gmx with the fepb200 hook (see README.md).  Needs /root/reference.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC=/tmp/gmxsrc
BUILD=/tmp/gmxbuild
OUT="$ROOT/integration/_gmx"
if [ ! -d "$SRC/src" ]; then
  rm -rf "$SRC"; mkdir -p "$SRC"; cp -r /root/reference/. "$SRC/"
fi
# (re-)apply the hook to a pristine copy of the one file it touches
cp /root/reference/src/gromacs/nbnxm/freeenergydispatch.cpp "$SRC/src/gromacs/nbnxm/freeenergydispatch.cpp"
# the source copy is shared with build_patched_gmx_cuda.sh, whose GPU-route patch touches three more files: this build
# takes them as they are in the reference
for f in src/gromacs/nbnxm/nbnxm_gpu_data_mgmt.cpp src/gromacs/nbnxm/cuda/nbnxm_cuda.cu src/gromacs/mdlib/sim_util.cpp; do
  cmp -s "/root/reference/$f" "$SRC/$f" || cp "/root/reference/$f" "$SRC/$f"
done
(cd "$SRC" && patch -p1 < "$ROOT/integration/gromacs_shim/freeenergydispatch_fepb200.patch")
# ... and the hook for the perturbed 1-4 pairs (listed_forces/pairs.cpp -> fepb200_pairs14_*)
cp /root/reference/src/gromacs/listed_forces/pairs.cpp "$SRC/src/gromacs/listed_forces/pairs.cpp"
(cd "$SRC" && patch -p1 < "$ROOT/integration/gromacs_shim/pairs_fepb200.patch")
# ... and the scope around the foreign-lambda loop of the listed forces (all points of the 1-4 pairs in one library call)
cp /root/reference/src/gromacs/listed_forces/listed_forces.cpp "$SRC/src/gromacs/listed_forces/listed_forces.cpp"
(cd "$SRC" && patch -p1 < "$ROOT/integration/gromacs_shim/listed_forces_fepb200.patch")
# ... and the hook for the non-perturbed cluster-pair kernel (nbnxm/kerneldispatch.cpp -> fepb200_nb_*)
cp /root/reference/src/gromacs/nbnxm/kerneldispatch.cpp "$SRC/src/gromacs/nbnxm/kerneldispatch.cpp"
(cd "$SRC" && patch -p1 < "$ROOT/integration/gromacs_shim/kerneldispatch_fepb200.patch")
mkdir -p "$BUILD"
cmake -G Ninja -S "$SRC" -B "$BUILD" -DCMAKE_C_COMPILER=/usr/bin/gcc -DCMAKE_CXX_COMPILER=/usr/bin/g++ \
  -DCMAKE_POLICY_VERSION_MINIMUM=3.5 -DGMX_GPU=OFF -DGMX_MPI=OFF -DGMX_THREAD_MPI=ON -DGMX_OPENMP=ON \
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
# the reference's coulandvdwtogether system as a slow-growth run: lambda moves from 0.5 by 0.005 per step, so every
# step needs the current lambda (the fork's GPU route uploads it once)
if [ ! -f "$TPR/coulandvdwtogether_slowgrowth.tpr" ]; then
  d="$SRC/src/testutils/simulationdatabase/freeenergy/coulandvdwtogether"
  sed 's/^init-lambda .*$/init-lambda              = 0.5\ndelta-lambda             = 0.005/' "$d/grompp.mdp" > /tmp/grompp_slowgrowth.mdp
  (cd /tmp && "$OUT/bin/gmx" -quiet grompp -f /tmp/grompp_slowgrowth.mdp -c "$d/conf.gro" -p "$d/topol.top" \
     -o "$TPR/coulandvdwtogether_slowgrowth.tpr" -po /tmp/mdout_slowgrowth.mdp -maxwarn 10 > /tmp/grompp_slowgrowth.log 2>&1) \
    || { echo "grompp failed for the slow-growth system"; tail -5 /tmp/grompp_slowgrowth.log; }
fi
ls -la "$OUT/bin" "$OUT/lib" "$TPR"
