/*
 * fep_list_build.cu -- builds the device layout of fep_types.h from the raw FEP t_nblist ON THE
 * GPU (search-step work).  Replaces, for this library, what the reference fork does on the host in
 * gpu_init_feppairlist() (nbnxm/nbnxm_gpu_data_mgmt.cpp:761-871: index remap + five H2D copies)
 * after combine_fep_lists() (nbnxm/pairlist.cpp:2867-2961), plus the preparation of the
 * atomic-free scatter (sorted destinations) that the fork does not have.
 *
 * The raw arrays are copied to the device once; everything else is kernels, prefix sums and stable
 * radix sorts (CUB):
 *   marks + scan           -> compact numbering of the touched atoms (ascending atom index)
 *   one thread per pair    -> 16-byte pair records (entry found by binary search in jindex),
 *                             segment-head flags
 *   scan of the head flags -> segment numbering, warp_hbase
 *   stable sort of {pairs by j atom, then segments by i atom} -> every force contribution's slot
 *                             in the atom-sorted buffer, atom_ptr
 *   stable sorts of the segments by shift index and by energy-group pair -> reduction ranges
 * All orders are the ones the host path of fepb200_set_list() produces (stable sorts keep the
 * slot / segment order inside one key), so both paths give bit-identical device structures.
 */
#include <cub/cub.cuh>

#include "fep_types.h"

namespace
{

__global__ void k_mark(const int* __restrict__ iinr, int nri, const int* __restrict__ jjnr, long long nrj,
                       int* __restrict__ mark)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nri)
    {
        mark[iinr[i]] = 1;
    }
    if (i < nrj)
    {
        mark[jjnr[i]] = 1;
    }
}

__global__ void k_touched(const int* __restrict__ mark, const int* __restrict__ cscan, int natoms, int* __restrict__ touched)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < natoms && mark[a])
    {
        touched[cscan[a]] = a;
    }
}

__global__ void k_entries(const int* __restrict__ iinr, const int* __restrict__ gid, const int* __restrict__ shift,
                          const int* __restrict__ cscan, int e0, int E, int4* __restrict__ ent4)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < E)
    {
        ent4[n] = make_int4(cscan[iinr[e0 + n]], shift[e0 + n], gid[e0 + n], 0);
    }
}

/* one thread per pair slot of this rank's shard */
__global__ void k_pairs(const int* __restrict__ jindex, const int* __restrict__ jjnr, const int* __restrict__ excl,
                        const int* __restrict__ cscan, const int4* __restrict__ ent4, int e0, int E, int j0, int P,
                        int4* __restrict__ pair4, int* __restrict__ keys, int* __restrict__ head)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P)
    {
        if (s == P)
        {
            head[P] = 0; /* terminator of the exclusive scan */
        }
        return;
    }
    const int k = j0 + s;
    /* the entry that holds pair k: last n with jindex[e0+n] <= k (empty entries are skipped) */
    int lo = 0, hi = E; /* invariant: jindex[e0+lo] <= k < jindex[e0+hi] */
    while (hi - lo > 1)
    {
        const int mid = (lo + hi) >> 1;
        if (jindex[e0 + mid] <= k)
        {
            lo = mid;
        }
        else
        {
            hi = mid;
        }
    }
    const int  n        = lo;
    const int4 en       = ent4[n];
    const int  cj       = cscan[jjnr[k]];
    const bool excluded = excl != nullptr && excl[k] == 0;
    pair4[s]            = make_int4(cj | (excluded ? (int)0x80000000u : 0), en.x | (en.y << 24), n, 0);
    keys[s]             = cj;
    head[s]             = ((s & 31) == 0 || k == jindex[e0 + n]) ? 1 : 0;
}

/* one thread per pair slot: the heads fill the per-segment arrays */
__global__ void k_segments(const int4* __restrict__ pair4, const int4* __restrict__ ent4, const int* __restrict__ head,
                           const int* __restrict__ hscan, int P, int* __restrict__ keys_seg /* = keys + P */,
                           int* __restrict__ seg_shift, int* __restrict__ seg_gid, int* __restrict__ warp_hbase)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P || !head[s])
    {
        return;
    }
    const int  h  = hscan[s];
    const int4 en = ent4[pair4[s].z];
    keys_seg[h]   = en.x;
    seg_shift[h]  = en.y;
    seg_gid[h]    = en.z;
    if ((s & 31) == 0)
    {
        warp_hbase[s >> 5] = h;
    }
}

__global__ void k_iota(int* __restrict__ v, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        v[i] = i;
    }
}

/* r-th element of the atom-sorted order: tell the contribution its slot, and fill atom_ptr at the
 * boundaries between different atoms (atoms without contributions get empty ranges) */
__global__ void k_atom_slots(const int* __restrict__ keys_sorted, const int* __restrict__ vals_sorted, int n, int P, int nT,
                             int4* __restrict__ pair4, int4* __restrict__ seg_dst, int* __restrict__ atom_ptr)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n)
    {
        return;
    }
    const int idx = vals_sorted[r];
    if (idx < P)
    {
        pair4[idx].w = r;
    }
    else
    {
        seg_dst[idx - P].x = r;
    }
    const int cur  = keys_sorted[r];
    const int prev = r > 0 ? keys_sorted[r - 1] : -1;
    for (int a = prev + 1; a <= cur; a++)
    {
        atom_ptr[a] = r;
    }
    if (r == n - 1)
    {
        for (int a = cur + 1; a <= nT; a++)
        {
            atom_ptr[a] = n;
        }
    }
}

/* r-th element of the segments sorted by `which` key (1: shift index -> seg_dst.y, 2: gid -> .z);
 * key_ptr[k] = first rank of key k, key_ptr[nkeys] = n */
__global__ void k_seg_slots(const int* __restrict__ keys_sorted, const int* __restrict__ vals_sorted, int n, int nkeys,
                            int which, int4* __restrict__ seg_dst, int* __restrict__ key_ptr)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n)
    {
        return;
    }
    const int h = vals_sorted[r];
    if (which == 1)
    {
        seg_dst[h].y = r;
    }
    else
    {
        seg_dst[h].z = r;
    }
    const int cur  = keys_sorted[r];
    const int prev = r > 0 ? keys_sorted[r - 1] : -1;
    for (int k = prev + 1; k <= cur; k++)
    {
        key_ptr[k] = r;
    }
    if (r == n - 1)
    {
        for (int k = cur + 1; k <= nkeys; k++)
        {
            key_ptr[k] = n;
        }
    }
}

int bits_for(int n)
{
    int b = 1;
    while ((1LL << b) < n && b < 31)
    {
        b++;
    }
    return b;
}

} // namespace

extern "C" size_t fep_list_build_temp_bytes(int natoms, long long n_sort_max)
{
    size_t a = 0, b = 0, c = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, a, (const int*)nullptr, (int*)nullptr, natoms + 1);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const int*)nullptr, (int*)nullptr, (int)n_sort_max + 1);
    cub::DeviceRadixSort::SortPairs(nullptr, c, (const int*)nullptr, (int*)nullptr, (const int*)nullptr, (int*)nullptr,
                                    (int)n_sort_max, 0, 31);
    return std::max(a, std::max(b, c)) + 256;
}

/* Phase 1: compact numbering.  mark/cscan: int[natoms+1]; returns after queuing (nT = cscan[natoms]). */
extern "C" int fep_list_build_touched(const int* d_iinr, int nri_total, const int* d_jjnr, long long nrj_total, int natoms,
                                      int* d_mark, int* d_cscan, int* d_touched, void* d_tmp, size_t tmp_bytes,
                                      cudaStream_t stream, long long* counter)
{
    cudaMemsetAsync(d_mark, 0, sizeof(int) * ((size_t)natoms + 1), stream);
    const long long n = std::max<long long>(nri_total, nrj_total);
    if (n > 0)
    {
        k_mark<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_iinr, nri_total, d_jjnr, nrj_total, d_mark);
        (*counter)++;
    }
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_mark, d_cscan, natoms + 1, stream);
    if (natoms > 0)
    {
        k_touched<<<(natoms + 255) / 256, 256, 0, stream>>>(d_mark, d_cscan, natoms, d_touched);
        (*counter)++;
    }
    return (int)cudaGetLastError();
}

/* Phase 2: pair records, head flags and their scan (H = hscan[P]). */
extern "C" int fep_list_build_pairs(const int* d_iinr, const int* d_gid, const int* d_shift, const int* d_jindex,
                                    const int* d_jjnr, const int* d_excl, const int* d_cscan, int e0, int E, int j0,
                                    int P, int4* d_ent4, int4* d_pair4, int* d_keys, int* d_head, int* d_hscan,
                                    void* d_tmp, size_t tmp_bytes, cudaStream_t stream, long long* counter)
{
    if (E > 0)
    {
        k_entries<<<(E + 255) / 256, 256, 0, stream>>>(d_iinr, d_gid, d_shift, d_cscan, e0, E, d_ent4);
        (*counter)++;
    }
    k_pairs<<<(P + 1 + 255) / 256, 256, 0, stream>>>(d_jindex, d_jjnr, d_excl, d_cscan, d_ent4, e0, E, j0, P, d_pair4,
                                                      d_keys, d_head);
    (*counter)++;
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_head, d_hscan, P + 1, stream);
    return (int)cudaGetLastError();
}

/* Phase 3 (H known): segments, the three stable sorts and the slot assignments.
 * keys: int[P+H] (first P filled by phase 2); scratch a/b: int[P+H] each for values, keys_out int[P+H];
 * seg_shift/seg_gid: int[H]; key_ptr: int[46 + G + 1] (shift_ptr then gid_ptr). */
extern "C" int fep_list_build_slots(const int4* d_ent4, int4* d_pair4, const int* d_head, const int* d_hscan, int P, int H,
                                    int nT, int ngrp, int* d_keys, int* d_keys_out, int* d_vals, int* d_vals_out,
                                    int* d_seg_shift, int* d_seg_gid, int* d_warp_hbase, int4* d_seg_dst, int* d_atom_ptr,
                                    int* d_key_ptr, void* d_tmp, size_t tmp_bytes, cudaStream_t stream,
                                    long long* counter)
{
    const int n = P + H;
    if (P > 0)
    {
        k_segments<<<(P + 255) / 256, 256, 0, stream>>>(d_pair4, d_ent4, d_head, d_hscan, P, d_keys + P, d_seg_shift,
                                                       d_seg_gid, d_warp_hbase);
        (*counter)++;
    }
    cudaMemsetAsync(d_atom_ptr, 0, sizeof(int) * ((size_t)nT + 1), stream);
    cudaMemsetAsync(d_key_ptr, 0, sizeof(int) * (FEP_NUM_SHIFT + 1 + ngrp + 1), stream);
    if (n > 0)
    {
        k_iota<<<(n + 255) / 256, 256, 0, stream>>>(d_vals, n);
        cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys_out, d_vals, d_vals_out, n, 0, bits_for(nT),
                                        stream);
        k_atom_slots<<<(n + 255) / 256, 256, 0, stream>>>(d_keys_out, d_vals_out, n, P, nT, d_pair4, d_seg_dst,
                                                         d_atom_ptr);
        (*counter) += 2;
    }
    if (H > 0)
    {
        k_iota<<<(H + 255) / 256, 256, 0, stream>>>(d_vals, H);
        cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_seg_shift, d_keys_out, d_vals, d_vals_out, H, 0, 6, stream);
        k_seg_slots<<<(H + 255) / 256, 256, 0, stream>>>(d_keys_out, d_vals_out, H, FEP_NUM_SHIFT, 1, d_seg_dst, d_key_ptr);
        cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_seg_gid, d_keys_out, d_vals, d_vals_out, H, 0,
                                        bits_for(ngrp), stream);
        k_seg_slots<<<(H + 255) / 256, 256, 0, stream>>>(d_keys_out, d_vals_out, H, ngrp, 2, d_seg_dst,
                                                        d_key_ptr + FEP_NUM_SHIFT + 1);
        (*counter) += 3;
    }
    return (int)cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------- */
/* peer exchange (fep_types.h): which rank PRODUCES every element of the atom-sorted contribution
 * buffer, of the shift-sorted segment forces and of the group-sorted segment energies, when rank r
 * evaluates the 32-pair warps [r * wpr, (r + 1) * wpr) of the flat pair space.  One byte per
 * element; the epilogue of the exchange reads each element from its producer's memory. */
__global__ void __launch_bounds__(256) k_source_tables(const int4* __restrict__ pair4, int P, const int4* __restrict__ seg_dst,
                                                       int H, const int* __restrict__ warp_hbase, int n_warps, int wpr,
                                                       unsigned char* __restrict__ slot_src,
                                                       unsigned char* __restrict__ fshift_src,
                                                       unsigned char* __restrict__ ev2_src)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < P)
    {
        slot_src[pair4[i].w] = (unsigned char)((i >> 5) / wpr);
    }
    if (i < n_warps)
    {
        /* the segments of warp i (at most 32) */
        const int           h0 = warp_hbase[i];
        const int           h1 = (i + 1 < n_warps) ? warp_hbase[i + 1] : H;
        const unsigned char r  = (unsigned char)(i / wpr);
        for (int h = h0; h < h1; h++)
        {
            const int4 sd    = seg_dst[h];
            slot_src[sd.x]   = r;
            fshift_src[sd.y] = r;
            ev2_src[sd.z]    = r;
        }
    }
}

extern "C" int fep_launch_source_tables(const int4* d_pair4, int P, const int4* d_seg_dst, int H, const int* d_warp_hbase,
                                        int wpr, unsigned char* d_slot_src, unsigned char* d_fshift_src,
                                        unsigned char* d_ev2_src, cudaStream_t stream, long long* counter)
{
    const int n_warps = (P + 31) / 32;
    if (P > 0 && wpr > 0)
    {
        k_source_tables<<<(P + 255) / 256, 256, 0, stream>>>(d_pair4, P, d_seg_dst, H, d_warp_hbase, n_warps, wpr, d_slot_src,
                                                            d_fshift_src, d_ev2_src);
        (*counter)++;
    }
    return (int)cudaGetLastError();
}
