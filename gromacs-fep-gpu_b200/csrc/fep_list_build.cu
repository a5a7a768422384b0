/*
 * fep_list_build.cu -- builds the device layout of fep_types.h from the raw FEP t_nblist ON THE
 * GPU (search-step work).  Replaces, for this library, what the reference fork does on the host in
 * gpu_init_feppairlist() (nbnxm/nbnxm_gpu_data_mgmt.cpp:761-871: index remap + five H2D copies)
 * after combine_fep_lists() (nbnxm/pairlist.cpp:2867-2961), plus the regrouping into trips and the
 * preparation of the atomic-free scatter (sorted destinations) that the fork does not have.
 *
 * The raw arrays are copied to the device once; everything else is kernels, prefix sums and stable
 * radix sorts (CUB):
 *   marks + scan            -> compact numbering of the touched atoms (ascending atom index)
 *   one thread per pair     -> both ends in compact numbering (entry found by binary search in
 *                              jindex), number of pairs every atom takes part in
 *   one thread per pair     -> owner = the end with more pairs; key = (owner, gid, shift, flipped)
 *   stable sort by key      -> groups; positions inside a group -> trips of <= 32 pairs -> slots
 *   one thread per pair     -> per-slot records (partner, pre-gathered partner parameters, the pair's
 *                              two type-table indices)
 *   one thread per trip     -> segments (runs of run_trips trips cut at group boundaries): the last
 *                              trip of a segment carries the owner's contribution
 *   stable sort of {pairs by partner atom, then segments by owner atom} -> every force contribution's
 *                              slot in the atom-sorted buffer, atom_ptr
 *   stable sorts of the segments by shift index and by energy-group pair -> reduction ranges
 * Stable sorts make the layout (and with it the order of every floating-point sum) a function of
 * the list alone.
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

/* one thread per pair of this rank's shard: its two ends, its entry, and the pair counts of both ends */
__global__ void k_pair_ends(const int* __restrict__ jindex, const int* __restrict__ jjnr, const int* __restrict__ cscan,
                            const int4* __restrict__ ent4, int e0, int E, int j0, int P, int* __restrict__ pj,
                            int* __restrict__ pn, int* __restrict__ deg)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P)
    {
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
    const int ci = ent4[lo].x;
    const int cj = cscan[jjnr[k]];
    pj[s]        = cj;
    pn[s]        = lo;
    atomicAdd(deg + ci, 1);
    atomicAdd(deg + cj, 1);
}

/* key = (((owner * G + gid) * 64 + shift) * 2 + flipped); the owner is the end that takes part in more
 * pairs (ties: the reference's i atom) */
template<typename K>
__global__ void k_pair_keys(const int* __restrict__ pj, const int* __restrict__ pn, const int4* __restrict__ ent4,
                            const int* __restrict__ deg, int G, int P, K* __restrict__ keys, int* __restrict__ vals)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P)
    {
        return;
    }
    const int4 en   = ent4[pn[s]];
    const int  ci   = en.x, cj = pj[s];
    const bool flip = deg[cj] > deg[ci];
    const K    og   = (K)(flip ? cj : ci) * (K)G + (K)en.z;
    keys[s]         = (og * 64 + (K)en.y) * 2 + (flip ? 1 : 0);
    vals[s]         = s;
}

template<typename K>
__global__ void k_group_marks(const K* __restrict__ keys_sorted, int P, int* __restrict__ gmark)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < P)
    {
        gmark[r] = (r == 0 || keys_sorted[r] != keys_sorted[r - 1]) ? r : 0;
    }
}

/* a trip starts every 32 pairs of a group */
__global__ void k_trip_heads(const int* __restrict__ gstart, int P, int* __restrict__ th)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < P)
    {
        th[r] = (((r - gstart[r]) & 31) == 0) ? 1 : 0;
    }
    else if (r == P)
    {
        th[P] = 0; /* terminator of the exclusive scan: tsc[P] = number of trips */
    }
}

__global__ void k_clear_trips(int NT, unsigned int* __restrict__ trips, int* __restrict__ orig)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x; /* one thread per lane slot */
    if (i < 32 * NT)
    {
        unsigned int* tb = trips + (size_t)(i >> 5) * FEP_TRIP_WORDS;
        const int     l  = i & 31;
        if (l < 16)
        {
            tb[l] = 0u;
        }
        tb[FEP_TW_CJX + l] = FEP_SLOT_PADDING;
        tb[FEP_TW_DST + l] = 0u;
        tb[FEP_TW_QA + l]  = 0u;
        tb[FEP_TW_QB + l]  = 0u;
        tb[FEP_TW_TJ + l]  = 0u;
        orig[i]            = -1;
    }
}

/* r-th pair of the group-sorted order -> its slot; the first pair of a trip also fills the trip's record */
template<typename K>
__global__ void k_fill_slots(const K* __restrict__ keys_sorted, const int* __restrict__ vals_sorted,
                             const int* __restrict__ gstart, const int* __restrict__ th, const int* __restrict__ tsc,
                             const int* __restrict__ pj, const int* __restrict__ pn, const int4* __restrict__ ent4,
                             const int* __restrict__ excl, int j0, const float4* __restrict__ par4, int ntype, int G, int P,
                             int run_mask, unsigned int* __restrict__ trips, int* __restrict__ orig, int* __restrict__ tshift,
                             int* __restrict__ tgid, int* __restrict__ tfirst, int* __restrict__ akeys,
                             int* __restrict__ avals)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P)
    {
        return;
    }
    const K    key    = keys_sorted[r];
    const int  s      = vals_sorted[r];
    const int  flip   = (int)(key & 1);
    const int  sh     = (int)((key >> 1) & 63);
    const K    og     = key >> 7;
    const int  owner  = (int)(og / (K)G);
    const int  g      = (int)(og - (K)owner * (K)G);
    const int  ci     = ent4[pn[s]].x, cj = pj[s];
    const int  other  = flip ? ci : cj;
    const int  t      = tsc[r] + th[r] - 1;
    const int  lane   = (r - gstart[r]) & 31;
    const bool excluded = excl != nullptr && excl[j0 + s] == 0;
    const float4  p   = par4[other];
    const float4  po  = par4[owner];
    unsigned int* tb  = trips + (size_t)t * FEP_TRIP_WORDS;
    tb[FEP_TW_CJX + lane] = (unsigned int)other | (excluded ? 0x80000000u : 0u);
    tb[FEP_TW_QA + lane]  = __float_as_uint(p.x);
    tb[FEP_TW_QB + lane]  = __float_as_uint(p.y);
    /* nbfp row = type of the reference's i atom (:499-500), column = type of its j atom (:560-563) */
    const int tiA = __float_as_int(flip ? p.z : po.z), tjA = __float_as_int(flip ? po.z : p.z);
    const int tiB = __float_as_int(flip ? p.w : po.w), tjB = __float_as_int(flip ? po.w : p.w);
    tb[FEP_TW_TJ + lane]  = (unsigned int)(ntype * tiA + tjA) | ((unsigned int)(ntype * tiB + tjB) << 16);
    orig[32 * t + lane]   = s;
    akeys[r]              = other;
    avals[r]              = 32 * t + lane;
    if (lane == 0)
    {
        const int sh_eff = flip ? (FEP_NUM_SHIFT - 1 - sh) : sh;
        tb[FEP_TH_OWNER] = (unsigned int)owner | ((unsigned int)sh_eff << 24) | (flip ? (unsigned int)FEP_TRIP_FLIPPED : 0u);
        tb[FEP_TH_QA]    = __float_as_uint(po.x);
        tb[FEP_TH_QB]    = __float_as_uint(po.y);
        tshift[t]        = sh;
        tgid[t]          = g;
        /* a segment starts with every run and with every group */
        tfirst[t]        = ((t & run_mask) == 0 || r == gstart[r]) ? 1 : 0;
    }
}

/* one thread per trip: the last trip of a segment carries the owner's contribution, the keys of the others sort
 * behind every real key (atom nT, shift 45, energy-group pair G) */
__global__ void k_segments(const int* __restrict__ tfirst, const int* __restrict__ tshift, const int* __restrict__ tgid, int NT,
                           int P, int nT, int G, unsigned int* __restrict__ trips, int* __restrict__ akeys,
                           int* __restrict__ avals, int* __restrict__ kshift, int* __restrict__ kgid)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= NT)
    {
        return;
    }
    const bool    last = t == NT - 1 || tfirst[t + 1] != 0;
    unsigned int* tb   = trips + (size_t)t * FEP_TRIP_WORDS;
    tb[FEP_TH_FLAGS]   = last ? FEP_TRIP_LAST : 0u;
    akeys[P + t]       = last ? (int)(tb[FEP_TH_OWNER] & (FEP_MAX_TOUCHED - 1)) : nT;
    avals[P + t]       = 32 * NT + t;
    kshift[t]          = last ? tshift[t] : FEP_NUM_SHIFT;
    kgid[t]            = last ? tgid[t] : G;
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
__global__ void k_atom_slots(const int* __restrict__ keys_sorted, const int* __restrict__ vals_sorted, int n, int n_slots,
                             int nT, unsigned int* __restrict__ trips, int* __restrict__ atom_ptr)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n)
    {
        return;
    }
    const int idx = vals_sorted[r];
    if (idx < n_slots)
    {
        trips[(size_t)(idx >> 5) * FEP_TRIP_WORDS + FEP_TW_DST + (idx & 31)] = (unsigned int)r;
    }
    else
    {
        trips[(size_t)(idx - n_slots) * FEP_TRIP_WORDS + FEP_TH_SLOT_F] = (unsigned int)r;
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

/* r-th element of the trips sorted by `which` key (1: shift index, 2: gid) -> the trip's slot in that order;
 * key_ptr[k] = first rank of key k, key_ptr[nkeys] = n */
__global__ void k_trip_slots(const int* __restrict__ keys_sorted, const int* __restrict__ vals_sorted, int n, int nkeys,
                             int which, unsigned int* __restrict__ trips, int* __restrict__ key_ptr)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n)
    {
        return;
    }
    const int t = vals_sorted[r];
    trips[(size_t)t * FEP_TRIP_WORDS + (which == 1 ? FEP_TH_SLOT_SHIFT : FEP_TH_SLOT_EV)] = (unsigned int)r;
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

int bits_for(unsigned long long n)
{
    int b = 1;
    while (b < 63 && (1ULL << b) < n)
    {
        b++;
    }
    return b;
}

struct MaxOp
{
    __device__ __forceinline__ int operator()(int a, int b) const { return a > b ? a : b; }
};

} // namespace

/* ---- hand-over of per-thread lists (fepb200_set_lists) ---------------------------------------- */
__global__ void k_add_offset(int* __restrict__ v, int n, int offset)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        v[i] += offset;
    }
}

/* atom indices of the lists -> indices of set_atoms through `map` (out-of-range entries become -1, which the range
 * check reports), then the range check: bad[0] counts the offending entries */
__global__ void k_remap_and_check(int* __restrict__ iinr, int nri, int* __restrict__ jjnr, long long nrj,
                                  const int* __restrict__ map, int n_map, int natoms, int* __restrict__ bad)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int which = 0; which < 2; which++)
    {
        int* v = which == 0 ? iinr : jjnr;
        if (i < (which == 0 ? (long long)nri : nrj))
        {
            int a = v[i];
            if (map != nullptr)
            {
                a    = (a >= 0 && a < n_map) ? map[a] : -1;
                v[i] = a;
            }
            if (a < 0 || a >= natoms)
            {
                atomicAdd(bad, 1);
            }
        }
    }
}

extern "C" int fep_list_add_offset(int* d_v, int n, int offset, cudaStream_t stream, long long* counter)
{
    if (n > 0 && offset != 0)
    {
        k_add_offset<<<(n + 255) / 256, 256, 0, stream>>>(d_v, n, offset);
        (*counter)++;
    }
    return (int)cudaGetLastError();
}

extern "C" int fep_list_remap_and_check(int* d_iinr, int nri, int* d_jjnr, long long nrj, const int* d_map, int n_map, int natoms,
                                        int* d_bad, cudaStream_t stream, long long* counter)
{
    cudaMemsetAsync(d_bad, 0, sizeof(int), stream);
    const long long n = std::max<long long>(nri, nrj);
    if (n > 0)
    {
        k_remap_and_check<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_iinr, nri, d_jjnr, nrj, d_map, n_map, natoms, d_bad);
        (*counter)++;
    }
    return (int)cudaGetLastError();
}

/* ---- the epilogue's atom records (fep_types.h): {atom, first contribution, one past the last, 0} of every atom
 * that receives contributions from this context's list, light and heavy atoms compacted into two arrays -------- */
struct RecIsLight
{
    __device__ __forceinline__ bool operator()(const int4& r) const
    {
        const int n = r.z - r.y;
        return n > 0 && n <= FEP_HEAVY_MIN;
    }
};
struct RecIsHeavy
{
    __device__ __forceinline__ bool operator()(const int4& r) const { return r.z - r.y > FEP_HEAVY_MIN; }
};

__global__ void k_atom_records(const int* __restrict__ atom_ptr, int nT, int4* __restrict__ rec)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < nT)
    {
        rec[a] = make_int4(a, atom_ptr[a], atom_ptr[a + 1], 0);
    }
}

/* d_counts[0] = light atoms, d_counts[1] = heavy atoms */
extern "C" int fep_list_build_records(const int* d_atom_ptr, int nT, int4* d_rec, int4* d_light, int4* d_heavy, int* d_counts,
                                      void* d_tmp, size_t tmp_bytes, cudaStream_t stream, long long* counter)
{
    cudaMemsetAsync(d_counts, 0, 2 * sizeof(int), stream);
    if (nT > 0)
    {
        k_atom_records<<<(nT + 255) / 256, 256, 0, stream>>>(d_atom_ptr, nT, d_rec);
        cub::DeviceSelect::If(d_tmp, tmp_bytes, d_rec, d_light, d_counts, nT, RecIsLight(), stream);
        cub::DeviceSelect::If(d_tmp, tmp_bytes, d_rec, d_heavy, d_counts + 1, nT, RecIsHeavy(), stream);
        (*counter) += 3;
    }
    return (int)cudaGetLastError();
}

extern "C" size_t fep_list_build_temp_bytes(int natoms, long long n_sort_max)
{
    size_t a = 0, b = 0, c = 0, d = 0, e = 0, f = 0;
    cub::DeviceSelect::If(nullptr, f, (const int4*)nullptr, (int4*)nullptr, (int*)nullptr, natoms + 1, RecIsLight());
    cub::DeviceScan::ExclusiveSum(nullptr, a, (const int*)nullptr, (int*)nullptr, natoms + 1);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const int*)nullptr, (int*)nullptr, (int)n_sort_max + 1);
    cub::DeviceRadixSort::SortPairs(nullptr, c, (const int*)nullptr, (int*)nullptr, (const int*)nullptr, (int*)nullptr,
                                    (int)n_sort_max, 0, 31);
    cub::DeviceRadixSort::SortPairs(nullptr, d, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (const int*)nullptr, (int*)nullptr, (int)n_sort_max, 0, 63);
    cub::DeviceScan::InclusiveScan(nullptr, e, (const int*)nullptr, (int*)nullptr, MaxOp(), (int)n_sort_max + 1);
    return std::max(std::max(std::max(a, e), f), std::max(b, std::max(c, d))) + 256;
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

template<typename K>
static void groups_typed(const ListBuild& b, int P, int G, int end_bit, cudaStream_t stream)
{
    K*     keys      = reinterpret_cast<K*>(b.keys);
    K*     keys_out  = reinterpret_cast<K*>(b.keys_out);
    size_t tmp_bytes = b.tmp_bytes;
    k_pair_keys<K><<<(P + 255) / 256, 256, 0, stream>>>(b.pj, b.pn, b.ent4, b.deg, G, P, keys, b.vals);
    cub::DeviceRadixSort::SortPairs(b.tmp, tmp_bytes, keys, keys_out, b.vals, b.vals_out, P, 0, end_bit, stream);
    k_group_marks<K><<<(P + 255) / 256, 256, 0, stream>>>(keys_out, P, b.gmark);
}

/* Phase 2: ends + pair counts, keys, the group sort and the trip numbering (NT = tsc[P]).
 * wide_keys: 64-bit sort keys (needed when nT * G * 128 does not fit 32 bits). */
extern "C" int fep_list_build_groups(const ListBuild* bp, const int* d_iinr, const int* d_gid, const int* d_shift,
                                     const int* d_jindex, const int* d_jjnr, const int* d_cscan, int e0, int E, int j0,
                                     int P, int nT, int G, int wide_keys, cudaStream_t stream, long long* counter)
{
    const ListBuild& b         = *bp;
    size_t           tmp_bytes = b.tmp_bytes;
    if (E > 0)
    {
        k_entries<<<(E + 255) / 256, 256, 0, stream>>>(d_iinr, d_gid, d_shift, d_cscan, e0, E, b.ent4);
        (*counter)++;
    }
    if (P <= 0)
    {
        cudaMemsetAsync(b.tsc, 0, sizeof(int), stream);
        return (int)cudaGetLastError();
    }
    cudaMemsetAsync(b.deg, 0, sizeof(int) * ((size_t)nT + 1), stream);
    k_pair_ends<<<(P + 255) / 256, 256, 0, stream>>>(d_jindex, d_jjnr, d_cscan, b.ent4, e0, E, j0, P, b.pj, b.pn, b.deg);
    const int end_bit = bits_for((unsigned long long)nT * (unsigned long long)G * 128ULL);
    if (wide_keys)
    {
        groups_typed<unsigned long long>(b, P, G, end_bit, stream);
    }
    else
    {
        groups_typed<unsigned int>(b, P, G, end_bit, stream);
    }
    cub::DeviceScan::InclusiveScan(b.tmp, tmp_bytes, b.gmark, b.gstart, MaxOp(), P, stream);
    k_trip_heads<<<(P + 1 + 255) / 256, 256, 0, stream>>>(b.gstart, P, b.th);
    cub::DeviceScan::ExclusiveSum(b.tmp, tmp_bytes, b.th, b.tsc, P + 1, stream);
    (*counter) += 4;
    return (int)cudaGetLastError();
}

/* Phase 3 (NT known): slot records, segments, the atom sort and the two segment sorts.  Can be repeated with another
 * run_trips (a power of two) as long as the results of phase 2 are in place.
 * key_ptr: int[46 + G + 1] (shift_ptr then gid_ptr); key_ptr[45] = key_ptr[46 + G] = atom_ptr[nT] - P = segments. */
extern "C" int fep_list_build_slots(const ListBuild* bp, const int* d_excl, int j0, const float4* d_par4, int ntype, int P,
                                    int NT, int nT, int G, int wide_keys, int run_trips, cudaStream_t stream,
                                    long long* counter)
{
    const ListBuild& b         = *bp;
    size_t           tmp_bytes = b.tmp_bytes;
    const int        n_slots   = 32 * NT;
    const int        n       = P + NT;
    cudaMemsetAsync(b.atom_ptr, 0, sizeof(int) * ((size_t)nT + 1), stream);
    cudaMemsetAsync(b.key_ptr, 0, sizeof(int) * (FEP_NUM_SHIFT + 1 + G + 1), stream);
    if (P <= 0)
    {
        return (int)cudaGetLastError();
    }
    k_clear_trips<<<(n_slots + 255) / 256, 256, 0, stream>>>(NT, b.trips, b.orig);
    if (wide_keys)
    {
        k_fill_slots<unsigned long long><<<(P + 255) / 256, 256, 0, stream>>>(
                reinterpret_cast<const unsigned long long*>(b.keys_out), b.vals_out, b.gstart, b.th, b.tsc, b.pj, b.pn, b.ent4,
                d_excl, j0, d_par4, ntype, G, P, run_trips - 1, b.trips, b.orig, b.tshift, b.tgid, b.tfirst, b.akeys, b.avals);
    }
    else
    {
        k_fill_slots<unsigned int><<<(P + 255) / 256, 256, 0, stream>>>(
                reinterpret_cast<const unsigned int*>(b.keys_out), b.vals_out, b.gstart, b.th, b.tsc, b.pj, b.pn, b.ent4, d_excl,
                j0, d_par4, ntype, G, P, run_trips - 1, b.trips, b.orig, b.tshift, b.tgid, b.tfirst, b.akeys, b.avals);
    }
    k_segments<<<(NT + 255) / 256, 256, 0, stream>>>(b.tfirst, b.tshift, b.tgid, NT, P, nT, G, b.trips, b.akeys, b.avals,
                                                     b.kshift, b.kgid);
    /* every force contribution's slot in the atom-sorted buffer: pairs (to their partner) in slot order, then
     * segments (to their owner); trips that do not end a segment sort behind the last atom */
    cub::DeviceRadixSort::SortPairs(b.tmp, tmp_bytes, b.akeys, b.akeys_out, b.avals, b.avals_out, n, 0,
                                    bits_for((unsigned long long)nT + 1), stream);
    k_atom_slots<<<(n + 255) / 256, 256, 0, stream>>>(b.akeys_out, b.avals_out, n, n_slots, nT, b.trips, b.atom_ptr);
    /* segments by shift index and by energy-group pair */
    k_iota<<<(NT + 255) / 256, 256, 0, stream>>>(b.avals, NT);
    cub::DeviceRadixSort::SortPairs(b.tmp, tmp_bytes, b.kshift, b.akeys_out, b.avals, b.avals_out, NT, 0, 6, stream);
    k_trip_slots<<<(NT + 255) / 256, 256, 0, stream>>>(b.akeys_out, b.avals_out, NT, FEP_NUM_SHIFT, 1, b.trips, b.key_ptr);
    cub::DeviceRadixSort::SortPairs(b.tmp, tmp_bytes, b.kgid, b.akeys_out, b.avals, b.avals_out, NT, 0,
                                    bits_for((unsigned long long)G + 1), stream);
    k_trip_slots<<<(NT + 255) / 256, 256, 0, stream>>>(b.akeys_out, b.avals_out, NT, G, 2, b.trips,
                                                     b.key_ptr + FEP_NUM_SHIFT + 1);
    (*counter) += 7;
    return (int)cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------- */
/* peer exchange (fep_types.h): which rank PRODUCES every element of the atom-sorted contribution
 * buffer, of the shift-sorted trip sums and of the group-sorted trip energies, when rank r
 * evaluates the trips [r * tpr, (r + 1) * tpr).  One byte per element; the epilogue of the
 * exchange reads each element from its producer's memory. */
__global__ void __launch_bounds__(256) k_source_tables(const unsigned int* __restrict__ trips, int NT, int tpr,
                                                       unsigned char* __restrict__ slot_src,
                                                       unsigned char* __restrict__ fshift_src,
                                                       unsigned char* __restrict__ ev2_src)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= 32 * NT)
    {
        return;
    }
    const int           t  = i >> 5, l = i & 31;
    const unsigned int* tb = trips + (size_t)t * FEP_TRIP_WORDS;
    const unsigned char r  = (unsigned char)(t / tpr);
    if (!(tb[FEP_TW_CJX + l] & FEP_SLOT_PADDING))
    {
        slot_src[tb[FEP_TW_DST + l]] = r;
    }
    if (l == 0 && (tb[FEP_TH_FLAGS] & FEP_TRIP_LAST))
    {
        slot_src[tb[FEP_TH_SLOT_F]]       = r;
        fshift_src[tb[FEP_TH_SLOT_SHIFT]] = r;
        ev2_src[tb[FEP_TH_SLOT_EV]]       = r;
    }
}

extern "C" int fep_launch_source_tables(const unsigned int* d_trips, int NT, int tpr, unsigned char* d_slot_src,
                                        unsigned char* d_fshift_src, unsigned char* d_ev2_src, cudaStream_t stream,
                                        long long* counter)
{
    if (NT > 0 && tpr > 0)
    {
        k_source_tables<<<(32 * NT + 255) / 256, 256, 0, stream>>>(d_trips, NT, tpr, d_slot_src, d_fshift_src, d_ev2_src);
        (*counter)++;
    }
    return (int)cudaGetLastError();
}
