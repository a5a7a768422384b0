/*
 * fepb200_api.cu -- host side of libfepb200.so: the C-ABI of include/fepb200.h.
 *
 * Owns the device-side layout (fep_types.h), turns the reference's inputs (t_nblist, per-atom
 * A/B charges and types, nbfp, interaction_const_t scalars, lambda vectors) into it, and drives
 * the kernels of fep_kernels.cu.  There is no CPU implementation of the pair mathematics in this
 * file or anywhere else in the library: without a CUDA device fepb200_create() fails.
 *
 * Reference call sites this replaces (paths under src/gromacs/):
 *   nbnxm/freeenergydispatch.cpp:147-308   dispatchFreeEnergyKernel() incl. foreign-lambda loop
 *   nbnxm/nbnxm_gpu_data_mgmt.cpp:491-536  cuda_copy_fepparams()
 *   nbnxm/nbnxm_gpu_data_mgmt.cpp:761-871  gpu_init_feppairlist()
 *   nbnxm/atomdata.cpp:1055-1072           setAtomPropertiesAB
 *   nbnxm/pairlist.cpp:2786-2838           balance_fep_lists() (here: split over ranks)
 */
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fepb200.h"
#include "fep_types.h"

/* fep_list_build.cu */
extern "C" size_t fep_list_build_temp_bytes(int natoms, long long n_sort_max);
extern "C" int    fep_list_build_touched(const int* d_iinr, int nri_total, const int* d_jjnr, long long nrj_total, int natoms,
                                         int* d_mark, int* d_cscan, int* d_touched, void* d_tmp, size_t tmp_bytes,
                                         cudaStream_t stream, long long* counter);
extern "C" int    fep_list_build_groups(const ListBuild* b, const int* d_iinr, const int* d_gid, const int* d_shift,
                                        const int* d_jindex, const int* d_jjnr, const int* d_cscan, int e0, int E, int j0, int P,
                                        int nT, int G, int wide_keys, cudaStream_t stream, long long* counter);
extern "C" int    fep_list_build_slots(const ListBuild* b, const int* d_excl, int j0, const float4* d_par4, int ntype, int P, int NT,
                                       int nT, int G, int wide_keys, int run_trips, cudaStream_t stream, long long* counter);
extern "C" int    fep_list_build_records(const int* d_atom_ptr, int nT, int4* d_rec, int4* d_light, int4* d_heavy, int* d_counts,
                                         void* d_tmp, size_t tmp_bytes, cudaStream_t stream, long long* counter);
extern "C" int    fep_list_add_offset(int* d_v, int n, int offset, cudaStream_t stream, long long* counter);
extern "C" int    fep_list_remap_and_check(int* d_iinr, int nri, int* d_jjnr, long long nrj, const int* d_map, int n_map, int natoms,
                                           int* d_bad, cudaStream_t stream, long long* counter);
extern "C" int    fep_launch_source_tables(const unsigned int* d_trips, int NT, int tpr, unsigned char* d_slot_src,
                                           unsigned char* d_fshift_src, unsigned char* d_ev2_src, cudaStream_t stream,
                                           long long* counter);

namespace
{

thread_local std::string g_create_error;

template<typename T>
struct DeviceArray
{
    T*     ptr = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap)
        {
            return cudaSuccess;
        }
        if (ptr)
        {
            cudaFree(ptr);
            ptr = nullptr;
            cap = 0;
        }
        const size_t want = n + n / 8 + 64;
        cudaError_t  e    = cudaMalloc(reinterpret_cast<void**>(&ptr), want * sizeof(T));
        if (e == cudaSuccess)
        {
            cap = want;
        }
        return e;
    }
    void release()
    {
        if (ptr)
        {
            cudaFree(ptr);
        }
        ptr = nullptr;
        cap = 0;
    }
};

template<typename T>
struct PinnedArray
{
    T*     ptr = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap)
        {
            return cudaSuccess;
        }
        if (ptr)
        {
            cudaFreeHost(ptr);
            ptr = nullptr;
            cap = 0;
        }
        const size_t want = n + n / 8 + 64;
        cudaError_t  e    = cudaMallocHost(reinterpret_cast<void**>(&ptr), want * sizeof(T));
        if (e == cudaSuccess)
        {
            cap = want;
        }
        return e;
    }
    void release()
    {
        if (ptr)
        {
            cudaFreeHost(ptr);
        }
        ptr = nullptr;
        cap = 0;
    }
};

constexpr int c_copy_chunks = 1; /* pipeline depth of the host<->device copies of large steps */

/* threads for the host-side gather / scatter of the touched atoms: about 4k atoms per thread, at
 * most 16 (measured on the 24-core B200 host, C5: gather 19 -> 9 us, scatter 18 -> 10 us against 8
 * threads of 16k atoms; more threads only add fork/join jitter) */
constexpr int c_host_grain = 4096;
int host_threads(int n)
{
    static const int cap  = std::getenv("FEPB200_HOST_THREADS") ? std::atoi(std::getenv("FEPB200_HOST_THREADS")) : 16;
    const int        want = std::max(1, n / c_host_grain);
    return std::min(std::min(want, std::max(1, cap)), std::max(1, omp_get_max_threads()));
}

float int_bits_as_float(int v)
{
    float f;
    std::memcpy(&f, &v, sizeof(f));
    return f;
}

bool eel_is_ewald(int e)
{
    /* md_enums.h: Pme 3, Ewald 4, P3mAD 5, PmeUser 13, PmeSwitch 14, PmeUserSwitch 15 */
    return e == 3 || e == 4 || e == 5 || e == 13 || e == 14 || e == 15;
}
bool eel_is_rf(int e)
{
    /* RF 1, GRF(removed) 2, RF_NEC 11, RFZero 16 */
    return e == 1 || e == 2 || e == 11 || e == 16;
}

} // namespace

struct fepb200_ctx
{
    int          device = -1;
    cudaStream_t stream = nullptr, own_stream = nullptr, side_stream = nullptr;
    /* FEPB200_TIMING: host wall-clock laps of compute() [gather, H2D queued, kernels queued, D2H queued,
     * results arrived, scatter done], printed by fepb200_destroy */
    bool         lap_on = false;
    double       lap_sum[6] = { 0, 0, 0, 0, 0, 0 };
    double       lap_t0 = 0;
    cudaEvent_t  lap_ev[4] = { nullptr, nullptr, nullptr, nullptr }; /* before/after H2D, after kernels, after D2H */
    double       lap_dev[3] = { 0, 0, 0 };
    long long    lap_n = 0;
    bool         chain_open = false; /* the last thing queued on `stream` is this step's epilogue (see fepb200_launch) */
    cudaEvent_t  fork_ev = nullptr, join_ev = nullptr;
    cudaEvent_t  ev_start = nullptr, ev_stop = nullptr;
    cudaEvent_t  ev_prof[4] = { nullptr, nullptr, nullptr, nullptr };
    cudaEvent_t  ev_copy[8] = {};
    bool         profiling = false, profiled = false;
    std::string  error;
    std::string  description;
    long long    launches = 0;
    bool         timed    = false;
    bool         staging_in_flight = false; /* an H2D copy out of h_step_in may still be running */
    /* ... and when that copy was queued by gather_x_device(): ev_staged marks its end, so the next gather waits for
     * the copy alone instead of draining a stream the caller shares with its own kernels (the nbnxm stream) */
    cudaEvent_t  ev_staged       = nullptr;
    cudaEvent_t  ev_handoff      = nullptr; /* fepb200_launch() on a stream of the caller's: orders it behind the context's stream */
    bool         staged_by_event = false;

    /* constants */
    bool           have_params = false;
    fepb200_params params{};
    int            softcore   = FEP_SC_NONE;
    int            elec_ewald = 0;
    int            foreign_mode = -1; /* >= 0: fused Beutler kernels of fep_beutler.cu */
    KernelArgs     ka{};

    /* nbfp */
    int                 ntype = 0;
    std::vector<float>  nbfp, nbfp_grid;
    DeviceArray<float4> d_typetab;
    bool                typetab_dirty = true;

    /* atoms */
    int                natoms = 0;
    std::vector<float> qA, qB;
    std::vector<int>   typeA, typeB;

    /* lambdas */
    bool                     have_lambda = false;
    float                    lam_c = 0, lam_v = 0;
    std::vector<float>       all_c, all_v;
    LambdaPoint              cur{};
    std::vector<LambdaPoint> pts;
    DeviceArray<LambdaPoint> d_pts;
    PinnedArray<LambdaPoint> h_pts;

    /* list */
    bool             have_list = false;
    fepb200_layout   layout{};
    int              first_entry = 0;
    std::vector<int> touched;    /* compact -> atom */
    /* fepb200_compute(): the epilogue writes the result block into the pinned host buffer itself */
    bool             zc_out = true, zc_next = false, result_on_host = false;
    unsigned char*   h_result_dev = nullptr; /* device view of the pinned result buffer (zero-copy output) */
    /* work arrays of fepb200_set_list(), kept between calls so that a search step does not pay for
     * page faults of fresh allocations */
    std::vector<int>    w_atom_ptr;
    std::vector<float4> w_par4;
    int              n_trips = 0;

    DeviceArray<int>    d_touched, d_atom_ptr, d_key_job_ptr;
    DeviceArray<int4>   d_ent4;
    DeviceArray<unsigned int> d_trips; /* [n_trips][FEP_TRIP_WORDS] */
    DeviceArray<int>    d_orig, d_tgid;
    DeviceArray<RedJob> d_red_jobs;
    DeviceArray<float4> d_par4, d_fsorted, d_fshift_sorted;
    DeviceArray<float2> d_ev2;
    DeviceArray<double> d_cta_part, d_for_part, d_job_part;
    DeviceArray<unsigned int> d_counter;
    /* raw list + scratch of the device-side list build (fep_list_build.cu) */
    DeviceArray<int> d_raw_iinr, d_raw_gid, d_raw_shift, d_raw_jindex, d_raw_jjnr, d_raw_excl, d_mark, d_cscan, d_pj, d_pn, d_deg,
            d_vals, d_vals_out, d_gmark, d_gstart, d_th, d_tsc, d_akeys, d_akeys_out, d_avals, d_avals_out, d_tshift, d_key_ptr;
    DeviceArray<int> d_tfirst, d_kshift, d_kgid, d_atom_map, d_bad;
    DeviceArray<unsigned long long> d_keys, d_keys_out;
    DeviceArray<unsigned char> d_cub_tmp;
    /* what phase 3 of the list build (slots + segments, build_segments()) needs to run again with another run length */
    struct
    {
        ListBuild  b{};
        const int* d_excl = nullptr;
        int        j0 = 0, P = 0, NT = 0, nT = 0, ngrp = 1, wide = 0;
    } lb;
    int n_segs = 0;
    DeviceArray<unsigned char> d_step_in; /* [DynHead | pos3[nT]] */
    DeviceArray<unsigned char> d_result;  /* [f64 block | f32 block] */
    PinnedArray<unsigned char> h_step_in, h_result;
    size_t res_f64_bytes = 0, res_f32_bytes = 0;
    unsigned char* res_target = nullptr; /* where the epilogue writes; nullptr = own result block */
    /* fepb200_set_push_targets(): the epilogue stores into the ranks' receive blocks (PushTargets) */
    int            push_n = 0;
    unsigned char* push_block[FEP_XMAX] = {};

    /* peer exchange (fepb200_set_peer_exchange): the list is evaluated by x_nranks GPUs, this one
     * takes a range of pairs and owns a range of atoms */
    bool           px_on = false;
    int            x_nranks = 1, x_rank = 0;
    int            x_range_trips = 0;   /* trips per rank, rounded up (same on every rank) */
    int            x_trip_begin = 0, x_trip_end = 0, x_atom_begin = 0, x_atom_end = 0, x_heavy_begin = 0, x_heavy_end = 0;
    DeviceArray<unsigned char> d_slot_src, d_fshift_src, d_ev2_src; /* producer rank of every sorted element */
    DeviceArray<int4>   d_heavy;      /* atoms with more than FEP_HEAVY_MIN force contributions, ascending: {atom, k0, k1, 0} */
    std::vector<int4>   heavy_atoms;  /* host copy */
    DeviceArray<int4>   d_light;      /* the other atoms with contributions from this context's list, ascending */
    std::vector<int4>   light_atoms;
    std::vector<int>    local_atoms;  /* compact atoms that occur in THIS context's pairs (= light + heavy), ascending: the
                                         only coordinates its kernels read */
    /* the records are built on the device; the host copies above (and w_atom_ptr) are fetched only by those who need
     * them: a shard of a split list (its gather) and the peer exchange (its atom ranges) */
    DeviceArray<int4>   d_atom_rec;
    DeviceArray<int>    d_rec_counts;
    bool                host_lists_valid = false;
    bool                shard_local      = false; /* upload_x gathers the atoms of local_atoms only */
    int                 x_light_begin = 0, x_light_end = 0;
    /* atoms of the compact numbering that receive nothing from this context's list are never written by the epilogue:
     * every result block is zeroed once per list (own blocks in prepare_buffers, a caller's block at its first step) */
    /* fepb200_reduce_scatter_peers(): after the reduction this context holds the forces of the compact atoms
     * [own_begin, own_end) only (zeros elsewhere) and all scalars */
    bool                own_on = false;
    int                 own_begin = 0, own_end = 0;
    bool                result_needs_zero = true;
    std::vector<void*>  zeroed_targets;
    DeviceArray<unsigned long long> d_trace;                        /* fepb200_epilogue_trace() */
    int                             trace_blocks = 0;               /* epilogue blocks of the last traced launch */
    unsigned int   x_seq = 0;
    unsigned char* x_base[FEP_XMAX] = {};
    size_t         x_bytes = 0; /* size of each rank's exchange buffer */
    size_t         x_off_fsorted = 0, x_off_fshift = 0, x_off_ev2 = 0, x_off_cta = 0, x_off_for = 0, x_slot_bytes = 0;
    /* why a kernel trapped (fep_fault in fep_types.h): pinned, mapped into the device, readable after the trap */
    unsigned int* h_fault = nullptr;
};

namespace
{

int fail(fepb200_ctx* ctx, int code, const char* fmt, ...)
{
    char    buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx)
    {
        ctx->error = buf;
        if (code == FEPB200_ERR_CUDA && ctx->h_fault && ctx->h_fault[0] != FEP_FAULT_NONE)
        {
            const unsigned int* w = ctx->h_fault;
            char                why[256];
            if (w[0] == FEP_FAULT_PEER_TIMEOUT)
            {
                snprintf(why, sizeof(why),
                         " [device: rank %u waited 4 s for rank %u to announce step %u of the peer exchange -- all ranks must "
                         "launch every step, in the same order]",
                         w[1], w[2], w[3]);
            }
            else if (w[0] == FEP_FAULT_STAGE_TIMEOUT)
            {
                snprintf(why, sizeof(why), " [device: the staged tile of CTA %u never arrived in shared memory (thread %u)]", w[1],
                         w[2]);
            }
            else
            {
                snprintf(why, sizeof(why), " [device fault %u: %u %u %u]", w[0], w[1], w[2], w[3]);
            }
            ctx->error += why;
        }
    }
    else
    {
        g_create_error = buf;
    }
    return code;
}

#define CU_CHECK(ctx, call)                                                                              \
    do                                                                                                   \
    {                                                                                                    \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
        {                                                                                                \
            return fail(ctx, FEPB200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                             \
        }                                                                                                \
    } while (0)

/* nb_free_energy.cpp:420-449 and :1405-1419, evaluated in double and rounded once */
LambdaPoint make_point(const fepb200_params& p, float lam_c_f, float lam_v_f)
{
    LambdaPoint  lp{};
    const double lc = lam_c_f, lv = lam_v_f;
    const double lfc[2] = { 1.0 - lc, lc }, lfv[2] = { 1.0 - lv, lv }, dl[2] = { -1.0, 1.0 };
    for (int s = 0; s < 2; s++)
    {
        const double oc = 1.0 - lfc[s], ov = 1.0 - lfv[s];
        const bool   p2 = p.lambdaPower == 2;
        lp.lfac_c[s]    = (float)lfc[s];
        lp.lfac_v[s]    = (float)lfv[s];
        lp.sclfac_c[s]  = (float)(p2 ? oc * oc : oc);
        lp.sclfac_v[s]  = (float)(p2 ? ov * ov : ov);
        lp.scdl_c[s]    = (float)(dl[s] * p.lambdaPower / 6.0 * (p2 ? oc : 1.0));
        lp.scdl_v[s]    = (float)(dl[s] * p.lambdaPower / 6.0 * (p2 ? ov : 1.0));
        lp.g6_c[s]      = (float)std::pow(oc, 1.0 / 6.0);
        lp.g6_v[s]      = (float)std::pow(ov, 1.0 / 6.0);
        lp.gdl_c[s]     = lfc[s] < 1.0 ? (float)(lfc[s] / oc) : 0.0f;
        lp.gdl_v[s]     = lfv[s] < 1.0 ? (float)(lfv[s] / ov) : 0.0f;
    }
    int differ = 1;
    if (p.alphaCoulomb == 0.0f && p.alphaVdw == 0.0f)
    {
        differ = 0;
    }
    else if (lam_c_f == lam_v_f && p.alphaCoulomb == p.alphaVdw)
    {
        differ = 0;
    }
    lp.differ = differ;
    return lp;
}

void refresh_points(fepb200_ctx* c)
{
    if (!c->have_params || !c->have_lambda)
    {
        return;
    }
    c->cur = make_point(c->params, c->lam_c, c->lam_v);
    c->pts.clear();
    c->pts.push_back(c->cur);
    for (size_t i = 0; i < c->all_c.size(); i++)
    {
        c->pts.push_back(make_point(c->params, c->all_c[i], c->all_v[i]));
    }
}

void fill_layout(fepb200_ctx* c)
{
    fepb200_layout& l  = c->layout;
    const long long g  = l.nenergrp;
    const long long np = l.nforeign + 1;
    l.f32_words        = 3LL * l.ntouched + 3 * FEP_NUM_SHIFT;
    l.off_fshift       = 3LL * l.ntouched;
    l.off_vc           = 0;
    l.off_vv           = g;
    l.off_dvdl         = 2 * g;
    l.off_foreign_e    = 2 * g + 2;
    l.off_foreign_dvdl = 2 * g + 2 + np;
    l.f64_words        = 2 * g + 2 + 3 * np;
}

/* (Re)allocates everything whose size depends on list, G and L, and refreshes the pointers
 * and sizes in the kernel argument block. */
int prepare_buffers(fepb200_ctx* c)
{
    fill_layout(c);
    const fepb200_layout& l = c->layout;
    KernelArgs&           k = c->ka;
    const int             np = l.nforeign + 1;

    k.n_points = np;
    int sms    = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    /* the trips this context evaluates: all of its list, or its share of it (peer exchange); every
     * launch size below follows from `range`, which is the same on all ranks of an exchange */
    const int range = c->px_on ? c->x_range_trips : k.n_trips;
    k.trip_begin    = c->px_on ? c->x_trip_begin : 0;
    k.trip_end      = c->px_on ? c->x_trip_end : k.n_trips;
    const int runs  = (range + k.run_trips - 1) / std::max(k.run_trips, 1);
    k.n_cta         = (runs + FEP_CTA / 32 - 1) / (FEP_CTA / 32); /* generic pass kernel: one warp per run */
    c->foreign_mode = -1;
    k.gapsys_hoisted = 0;
    if (c->softcore == FEP_SC_BEUTLER && !k.pot_switch)
    {
        /* nb_free_energy.cpp:1405-1419 decides per call whether the Coulomb and LJ soft-core radii
         * differ; here per set of lambda points */
        if (c->params.alphaCoulomb == 0.0f)
        {
            c->foreign_mode = 0;
        }
        else
        {
            bool same = c->params.alphaCoulomb == c->params.alphaVdw && c->lam_c == c->lam_v;
            for (size_t i = 0; i < c->all_c.size(); i++)
            {
                same = same && c->all_c[i] == c->all_v[i];
            }
            c->foreign_mode = same ? 1 : 2;
        }
        /* one launch per chunk; split the points only when the trips alone cannot fill the GPU */
        const long long pair_ctas = (range + FEP_FB_CTA / 32 - 1) / (FEP_FB_CTA / 32);
        int             want      = 1;
        if (pair_ctas > 0 && pair_ctas < 2LL * sms)
        {
            want = (int)std::min<long long>(np, (2LL * sms + pair_ctas - 1) / pair_ctas);
        }
        k.chunk_points = fep_beutler_chunk_size(np, want);
        k.n_chunks     = (np + k.chunk_points - 1) / k.chunk_points;
        /* a warp evaluates whole runs, run after run with a grid-wide stride, warp-major (the first warps of all CTAs
         * first): one full wave of resident CTAs whenever there is a run for every CTA, so that the busy warps are
         * spread evenly over the SMs; fewer CTAs only for lists with fewer runs than that */
        auto tiles = [&](long long ctas_per_sm, bool with_force, int& n_tiles) {
            const long long units = with_force ? runs : range; /* energy-only launches walk single trips */
            n_tiles               = units > 0 ? (int)std::max(1LL, std::min(units, (long long)sms * ctas_per_sm)) : 0;
        };
        /* fuse pass + foreign when the list is too small to fill the GPU anyway */
        k.fuse_pass_and_foreign = pair_ctas < 16LL * sms;
        if (const char* env = std::getenv("FEPB200_FUSE"))
        {
            k.fuse_pass_and_foreign = std::atoi(env) != 0;
        }
        /* occupancy of the very kernels the step will launch (cudaOccupancyMaxActiveBlocksPerMultiprocessor).
         * FEPB200_FOREIGN_CTAS_PER_SM / FEPB200_PASS_CTAS_PER_SM size the grids for fewer resident CTAs
         * than fit (experiment: a pass grid and a foreign grid that are co-resident on every SM and
         * fill each other's stalls, instead of one full wave after the other) */
        auto per_sm = [](const char* env, int occ) {
            const char* e = std::getenv(env);
            const int   v = e ? std::atoi(e) : 0;
            return v > 0 ? std::min(v, occ) : occ;
        };
        tiles(per_sm("FEPB200_FOREIGN_CTAS_PER_SM",
                     fep_beutler_ctas_per_sm(c->elec_ewald, c->foreign_mode, k.chunk_points, k.fuse_pass_and_foreign)),
              k.fuse_pass_and_foreign != 0, k.n_tiles);
        tiles(per_sm("FEPB200_PASS_CTAS_PER_SM", fep_beutler_ctas_per_sm(c->elec_ewald, c->foreign_mode, 0, 1)), true,
              k.pass_n_tiles);
    }
    else if (c->softcore == FEP_SC_GAPSYS && !k.pot_switch && std::getenv("FEPB200_GAPSYS_GENERIC") == nullptr)
    {
        /* fep_gapsys.cu: single trips dealt to the warps of one wave of CTAs, chunks of at most 8 points */
        k.gapsys_hoisted = 1;
        k.chunk_points   = fep_gapsys_chunk_size(np);
        k.n_chunks       = (np + k.chunk_points - 1) / k.chunk_points;
        k.tile_trips     = 0;
        /* grid = n_tiles x n_chunks CTAs: about one wave in total */
        const long long wave = (long long)sms * fep_gapsys_ctas_per_sm(c->elec_ewald, k.chunk_points);
        k.n_tiles            = range > 0 ? (int)std::max(1LL, std::min<long long>(range, (wave + k.n_chunks - 1) / k.n_chunks)) : 0;
    }
    else
    {
        /* generic kernel: lambda chunks of at most FEP_LCHUNK points, evenly sized */
        k.n_chunks     = (np + FEP_LCHUNK - 1) / FEP_LCHUNK;
        k.chunk_points = (np + k.n_chunks - 1) / k.n_chunks;
        k.n_chunks     = (np + k.chunk_points - 1) / k.chunk_points;
        /* tiles of trips: enough CTAs to fill the GPU several times over, but several trips per warp
         * on large lists to amortise the final reduction */
        const long long target_ctas = 8LL * sms;
        long long       per_warp    = ((long long)range * k.n_chunks + target_ctas * (FEP_CTA / 32) - 1)
                             / (target_ctas * (FEP_CTA / 32));
        per_warp     = std::max(1LL, std::min(per_warp, 8LL));
        k.tile_trips = (int)per_warp * (FEP_CTA / 32);
        k.n_tiles    = (range + k.tile_trips - 1) / k.tile_trips;
    }
    if (c->px_on)
    {
        /* carve one exchange slot: [fsorted | fshift_sorted | ev2 | cta_part | for_part], every part
         * 256-byte aligned; identical on all ranks because every input of the sizes is */
        auto         up      = [](size_t b) { return (b + 255) & ~(size_t)255; };
        const size_t P       = (size_t)k.n_pairs, H = (size_t)k.n_segs;
        const size_t n_parts = (size_t)std::max(std::max(k.n_cta, k.n_tiles), std::max(k.pass_n_tiles * (FEP_FB_CTA / 32), 1));
        c->x_off_fsorted     = 0;
        c->x_off_fshift      = c->x_off_fsorted + up((P + H) * sizeof(float4));
        c->x_off_ev2         = c->x_off_fshift + up(H * sizeof(float4));
        c->x_off_cta         = c->x_off_ev2 + up(H * sizeof(float2));
        c->x_off_for         = c->x_off_cta + up(4 * n_parts * sizeof(double));
        c->x_slot_bytes      = c->x_off_for + up(3 * (size_t)np * (size_t)std::max(k.n_tiles, 1) * sizeof(double));
        if (2 * c->x_slot_bytes + 256 > c->x_bytes)
        {
            return fail(c, FEPB200_ERR_STATE,
                        "peer exchange buffers are too small for this list / number of lambda points (%zu bytes needed, "
                        "%zu given): call fepb200_set_peer_exchange() again",
                        2 * c->x_slot_bytes + 256, c->x_bytes);
        }
    }

    CU_CHECK(c, c->d_pts.reserve(np));
    CU_CHECK(c, c->h_pts.reserve(np));
    CU_CHECK(c, c->d_cta_part.reserve(
                        4 * (size_t)std::max(std::max(k.n_cta, k.n_tiles), std::max(k.pass_n_tiles * (FEP_FB_CTA / 32), 1))));
    CU_CHECK(c, c->d_for_part.reserve(3 * (size_t)np * std::max(k.n_tiles, 1)));
    c->res_f64_bytes = ((size_t)l.f64_words * sizeof(double) + 15) & ~(size_t)15;
    c->res_f32_bytes = (size_t)l.f32_words * sizeof(float);
    {
        const unsigned char* before_d = c->d_result.ptr;
        const unsigned char* before_h = c->h_result.ptr;
        CU_CHECK(c, c->d_result.reserve(c->res_f64_bytes + c->res_f32_bytes));
        CU_CHECK(c, c->h_result.reserve(c->res_f64_bytes + c->res_f32_bytes));
        if (c->result_needs_zero || before_d != c->d_result.ptr || before_h != c->h_result.ptr)
        {
            CU_CHECK(c, cudaStreamSynchronize(c->stream)); /* nobody reads or writes the blocks any more */
            CU_CHECK(c, cudaMemsetAsync(c->d_result.ptr, 0, c->res_f64_bytes + c->res_f32_bytes, c->stream));
            std::memset(c->h_result.ptr, 0, c->res_f64_bytes + c->res_f32_bytes);
            c->result_needs_zero = false;
        }
    }
    {
        void* dp = nullptr;
        c->h_result_dev =
                (cudaHostGetDevicePointer(&dp, c->h_result.ptr, 0) == cudaSuccess) ? static_cast<unsigned char*>(dp) : nullptr;
        cudaGetLastError();
    }
    const size_t step_bytes = sizeof(DynHead) + 3 * sizeof(float) * (size_t)l.ntouched;
    CU_CHECK(c, c->d_step_in.reserve(step_bytes));
    CU_CHECK(c, c->h_step_in.reserve(step_bytes));

    k.dyn      = reinterpret_cast<const DynHead*>(c->d_step_in.ptr);
    k.pos3     = reinterpret_cast<const float*>(c->d_step_in.ptr + sizeof(DynHead));
    k.pts      = c->d_pts.ptr;
    k.cta_part = c->d_cta_part.ptr;
    k.for_part = c->d_for_part.ptr;
    k.res_f64  = reinterpret_cast<double*>(c->d_result.ptr);
    k.res_f32  = reinterpret_cast<float*>(c->d_result.ptr + c->res_f64_bytes);
    c->res_target = nullptr;
    return FEPB200_OK;
}

int upload_points(fepb200_ctx* c)
{
    if (c->pts.empty())
    {
        return FEPB200_OK;
    }
    CU_CHECK(c, c->d_pts.reserve(c->pts.size()));
    CU_CHECK(c, c->h_pts.reserve(c->pts.size()));
    c->ka.pts = c->d_pts.ptr;
    /* the previous copy out of the pinned buffer must have completed */
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    std::memcpy(c->h_pts.ptr, c->pts.data(), c->pts.size() * sizeof(LambdaPoint));
    CU_CHECK(c, cudaMemcpyAsync(c->d_pts.ptr, c->h_pts.ptr, c->pts.size() * sizeof(LambdaPoint),
                                cudaMemcpyHostToDevice, c->stream));
    return FEPB200_OK;
}

int upload_typetab(fepb200_ctx* c)
{
    if (!c->typetab_dirty || c->ntype == 0 || !c->have_params)
    {
        return FEPB200_OK;
    }
    const int           t = c->ntype;
    std::vector<float4> tab((size_t)t * t);
    const bool          gapsys = c->params.softcoreType == FEPB200_SC_GAPSYS;
    for (int i = 0; i < t * t; i++)
    {
        const float c6 = c->nbfp[2 * i], c12 = c->nbfp[2 * i + 1];
        float       sig6;
        /* nb_free_energy.cpp:571-594; the division is done in fp32 like the mixed-precision build */
        if (c6 > 0.0f && c12 > 0.0f)
        {
            sig6 = 0.5f * c12 / c6;
            if (!gapsys && sig6 < c->params.sigma6Minimum)
            {
                sig6 = c->params.sigma6Minimum;
            }
        }
        else
        {
            sig6 = gapsys ? c->params.gapsysSigma6VdW : c->params.sigma6WithInvalidSigma;
        }
        const float c6g = c->nbfp_grid.empty() ? 0.0f : c->nbfp_grid[2 * i];
        tab[i]          = make_float4(c6, c12, sig6, c6g);
    }
    CU_CHECK(c, c->d_typetab.reserve(tab.size()));
    CU_CHECK(c, cudaMemcpyAsync(c->d_typetab.ptr, tab.data(), tab.size() * sizeof(float4), cudaMemcpyHostToDevice,
                                c->stream));
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    c->ka.typetab    = c->d_typetab.ptr;
    c->ka.ntype      = t;
    c->typetab_dirty = false;
    return FEPB200_OK;
}

template<typename T>
int to_device(fepb200_ctx* c, DeviceArray<T>& d, const std::vector<T>& h)
{
    CU_CHECK(c, d.reserve(std::max<size_t>(h.size(), 1)));
    if (!h.empty())
    {
        CU_CHECK(c, cudaMemcpyAsync(d.ptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    }
    return FEPB200_OK;
}

StepFlags step_flags(const fepb200_ctx* c, int flags)
{
    StepFlags sf;
    sf.force   = (flags & FEPB200_DO_FORCE) != 0;
    sf.shift   = sf.force && (flags & FEPB200_DO_SHIFTFORCE) != 0; /* nb_free_energy.cpp:1153-1164 */
    sf.energy  = (flags & FEPB200_DO_POTENTIAL) != 0;
    sf.foreign = (flags & FEPB200_DO_FOREIGNLAMBDA) != 0 && !c->pts.empty();
    return sf;
}

int check_ready(fepb200_ctx* c)
{
    if (!c->have_params)
    {
        return fail(c, FEPB200_ERR_STATE, "fepb200_set_params() has not been called");
    }
    if (c->ntype == 0)
    {
        return fail(c, FEPB200_ERR_STATE, "fepb200_set_nbfp() has not been called");
    }
    if (!c->have_list)
    {
        return fail(c, FEPB200_ERR_STATE, "fepb200_set_list() has not been called");
    }
    if (!c->have_lambda)
    {
        return fail(c, FEPB200_ERR_STATE, "fepb200_set_lambdas() has not been called");
    }
    if (c->params.vdwtype == FEPB200_VDW_PME && c->nbfp_grid.empty())
    {
        return fail(c, FEPB200_ERR_STATE, "LJ-PME needs nbfp_grid in fepb200_set_nbfp()");
    }
    return upload_typetab(c);
}

} // namespace

/* Reduction jobs of the epilogue from the prefix arrays of the shift-sorted and gid-sorted
 * segment orders: chunks of at most FEP_RED_CHUNK elements per key, shift keys first. */
static void build_jobs(const int* shift_ptr /*[46]*/, const int* gid_ptr /*[G+1]*/, int ngrp, std::vector<RedJob>* jobs,
                       std::vector<int>* key_job_ptr, int* n_shift_jobs)
{
    jobs->clear();
    key_job_ptr->assign(FEP_NUM_SHIFT + ngrp + 1, 0);
    for (int kind = 0; kind < 2; kind++)
    {
        const int  nkeys = kind == 0 ? FEP_NUM_SHIFT : ngrp;
        const int* ptr   = kind == 0 ? shift_ptr : gid_ptr;
        for (int k = 0; k < nkeys; k++)
        {
            (*key_job_ptr)[(kind == 0 ? 0 : FEP_NUM_SHIFT) + k] = (int)jobs->size();
            for (int b = ptr[k]; b < ptr[k + 1]; b += FEP_RED_CHUNK)
            {
                RedJob j;
                j.begin = b;
                j.end   = std::min(ptr[k + 1], b + FEP_RED_CHUNK);
                j.key   = k;
                j.kind  = kind;
                jobs->push_back(j);
            }
        }
        if (kind == 0)
        {
            *n_shift_jobs = (int)jobs->size();
        }
    }
    (*key_job_ptr)[FEP_NUM_SHIFT + ngrp] = (int)jobs->size();
}

/* The device-side builder of the list layout (fep_list_build.cu): raw list to the GPU, then
 * kernels, scans and stable sorts; only two scalars (nT, H), the touched-atom list and 47 + G
 * counters come back to the host. */
static int build_list_device(fepb200_ctx* c, int n_lists, const fepb200_list_view* lists, const int* atom_map, int n_map, int nri,
                             long long nrj_total, int ngrp, int e0, int E, int j0, int P)
{
    cudaStream_t st = c->stream;
    /* raw lists: every list's arrays go straight to their place in the concatenated device arrays */
    CU_CHECK(c, c->d_raw_iinr.reserve(std::max(nri, 1)));
    CU_CHECK(c, c->d_raw_gid.reserve(std::max(nri, 1)));
    CU_CHECK(c, c->d_raw_shift.reserve(std::max(nri, 1)));
    CU_CHECK(c, c->d_raw_jindex.reserve((size_t)nri + 1));
    CU_CHECK(c, c->d_raw_jjnr.reserve(std::max<long long>(nrj_total, 1)));
    CU_CHECK(c, c->d_bad.reserve(1));
    const bool have_excl = n_lists > 0 && lists[0].excl_fep != nullptr && nrj_total > 0;
    if (have_excl)
    {
        CU_CHECK(c, c->d_raw_excl.reserve(nrj_total));
    }
    {
        int       eo = 0;
        long long po = 0;
        for (int l = 0; l < n_lists; l++)
        {
            const fepb200_list_view& v = lists[l];
            const long long          np = v.nri > 0 ? v.jindex[v.nri] : 0;
            if (v.nri > 0)
            {
                CU_CHECK(c, cudaMemcpyAsync(c->d_raw_iinr.ptr + eo, v.iinr, sizeof(int) * v.nri, cudaMemcpyHostToDevice, st));
                CU_CHECK(c, cudaMemcpyAsync(c->d_raw_gid.ptr + eo, v.gid, sizeof(int) * v.nri, cudaMemcpyHostToDevice, st));
                CU_CHECK(c, cudaMemcpyAsync(c->d_raw_shift.ptr + eo, v.shift, sizeof(int) * v.nri, cudaMemcpyHostToDevice, st));
                CU_CHECK(c, cudaMemcpyAsync(c->d_raw_jindex.ptr + eo, v.jindex, sizeof(int) * v.nri, cudaMemcpyHostToDevice, st));
                const int err = fep_list_add_offset(c->d_raw_jindex.ptr + eo, v.nri, (int)po, st, &c->launches);
                if (err != 0)
                {
                    return fail(c, FEPB200_ERR_CUDA, "list hand-over failed: %s", cudaGetErrorString((cudaError_t)err));
                }
            }
            if (np > 0)
            {
                CU_CHECK(c, cudaMemcpyAsync(c->d_raw_jjnr.ptr + po, v.jjnr, sizeof(int) * np, cudaMemcpyHostToDevice, st));
                if (have_excl)
                {
                    CU_CHECK(c, cudaMemcpyAsync(c->d_raw_excl.ptr + po, v.excl_fep, sizeof(int) * np, cudaMemcpyHostToDevice, st));
                }
            }
            eo += v.nri;
            po += np;
        }
        const int last = (int)nrj_total;
        CU_CHECK(c, cudaMemcpyAsync(c->d_raw_jindex.ptr + nri, &last, sizeof(int), cudaMemcpyHostToDevice, st));
    }
    const int* d_excl = have_excl ? c->d_raw_excl.ptr : nullptr;
    /* index space of the lists -> index space of set_atoms, and the range check of every atom index, on the device */
    const int* d_map = nullptr;
    if (atom_map != nullptr && n_map > 0)
    {
        CU_CHECK(c, c->d_atom_map.reserve(n_map));
        CU_CHECK(c, cudaMemcpyAsync(c->d_atom_map.ptr, atom_map, sizeof(int) * (size_t)n_map, cudaMemcpyHostToDevice, st));
        d_map = c->d_atom_map.ptr;
    }
    {
        const int err = fep_list_remap_and_check(c->d_raw_iinr.ptr, nri, c->d_raw_jjnr.ptr, nrj_total, d_map, n_map, c->natoms,
                                                 c->d_bad.ptr, st, &c->launches);
        if (err != 0)
        {
            return fail(c, FEPB200_ERR_CUDA, "list hand-over failed: %s", cudaGetErrorString((cudaError_t)err));
        }
        int bad = 0;
        CU_CHECK(c, cudaMemcpyAsync(&bad, c->d_bad.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU_CHECK(c, cudaStreamSynchronize(st));
        if (bad != 0)
        {
            return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_set_list: %d atom indices of the list%s are outside [0,%d)", bad,
                        d_map ? " (after the atom map)" : "", c->natoms);
        }
    }
    if (c->ntype > FEP_MAX_NTYPE)
    {
        return fail(c, FEPB200_ERR_UNSUPPORTED, "more than %d atom types", FEP_MAX_NTYPE);
    }
    /* compact numbering of the atoms of the FULL list */
    const long long n_max     = 2LL * P + 64; /* P pairs + at most P trips */
    if (n_max >= (1LL << 31) - 64 || 33LL * P >= (1LL << 31) - 64)
    {
        return fail(c, FEPB200_ERR_UNSUPPORTED, "list too large");
    }
    const size_t tmp_bytes = fep_list_build_temp_bytes(c->natoms, n_max);
    CU_CHECK(c, c->d_cub_tmp.reserve(tmp_bytes));
    CU_CHECK(c, c->d_mark.reserve((size_t)c->natoms + 1));
    CU_CHECK(c, c->d_cscan.reserve((size_t)c->natoms + 1));
    CU_CHECK(c, c->d_touched.reserve(std::max(c->natoms, 1)));
    int err = fep_list_build_touched(c->d_raw_iinr.ptr, nri, c->d_raw_jjnr.ptr, nrj_total, c->natoms, c->d_mark.ptr,
                                     c->d_cscan.ptr, c->d_touched.ptr, c->d_cub_tmp.ptr, tmp_bytes, st, &c->launches);
    if (err != 0)
    {
        return fail(c, FEPB200_ERR_CUDA, "list build (touched) failed: %s", cudaGetErrorString((cudaError_t)err));
    }
    int nT = 0, NT = 0;
    CU_CHECK(c, cudaMemcpyAsync(&nT, c->d_cscan.ptr + c->natoms, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_CHECK(c, cudaStreamSynchronize(st));
    if (nT >= FEP_MAX_TOUCHED)
    {
        return fail(c, FEPB200_ERR_UNSUPPORTED, "more than %d distinct atoms in one FEP list", FEP_MAX_TOUCHED);
    }
    /* scratch and the buffers whose size follows from P and nT; the group sort does not need the atom parameters,
     * so it is queued before the host prepares them */
    const bool wide = (unsigned long long)std::max(nT, 1) * (unsigned long long)ngrp * 128ULL >= (1ULL << 32);
    CU_CHECK(c, c->d_ent4.reserve(std::max(E, 1)));
    CU_CHECK(c, c->d_pj.reserve(std::max(P, 1)));
    CU_CHECK(c, c->d_pn.reserve(std::max(P, 1)));
    CU_CHECK(c, c->d_deg.reserve((size_t)nT + 1));
    CU_CHECK(c, c->d_keys.reserve(std::max(P, 1)));
    CU_CHECK(c, c->d_keys_out.reserve(std::max(P, 1)));
    CU_CHECK(c, c->d_vals.reserve(std::max(P, 1)));
    CU_CHECK(c, c->d_vals_out.reserve(std::max(P, 1)));
    CU_CHECK(c, c->d_gmark.reserve(std::max(P, 1)));
    CU_CHECK(c, c->d_gstart.reserve(std::max(P, 1)));
    CU_CHECK(c, c->d_th.reserve((size_t)P + 1));
    CU_CHECK(c, c->d_tsc.reserve((size_t)P + 1));
    CU_CHECK(c, c->d_atom_ptr.reserve((size_t)nT + 1));
    CU_CHECK(c, c->d_key_ptr.reserve(FEP_NUM_SHIFT + 1 + ngrp + 1));
    ListBuild b{};
    b.pj = c->d_pj.ptr, b.pn = c->d_pn.ptr, b.deg = c->d_deg.ptr;
    b.keys = c->d_keys.ptr, b.keys_out = c->d_keys_out.ptr, b.vals = c->d_vals.ptr, b.vals_out = c->d_vals_out.ptr;
    b.gmark = c->d_gmark.ptr, b.gstart = c->d_gstart.ptr, b.th = c->d_th.ptr, b.tsc = c->d_tsc.ptr;
    b.tmp = c->d_cub_tmp.ptr, b.tmp_bytes = tmp_bytes;
    b.ent4 = c->d_ent4.ptr, b.atom_ptr = c->d_atom_ptr.ptr, b.key_ptr = c->d_key_ptr.ptr;
    err = fep_list_build_groups(&b, c->d_raw_iinr.ptr, c->d_raw_gid.ptr, c->d_raw_shift.ptr, c->d_raw_jindex.ptr,
                                c->d_raw_jjnr.ptr, c->d_cscan.ptr, e0, E, j0, P, nT, ngrp, wide ? 1 : 0, st, &c->launches);
    if (err != 0)
    {
        return fail(c, FEPB200_ERR_CUDA, "list build (groups) failed: %s", cudaGetErrorString((cudaError_t)err));
    }
    CU_CHECK(c, cudaMemcpyAsync(&NT, c->d_tsc.ptr + P, sizeof(int), cudaMemcpyDeviceToHost, st));
    /* while the GPU sorts: the touched atoms (for the coordinate gather / force scatter of every step) and the
     * per-atom parameters in compact order (the per-atom arrays of set_atoms live on the host) */
    c->touched.resize(nT);
    if (nT > 0)
    {
        CU_CHECK(c, cudaMemcpyAsync(c->touched.data(), c->d_touched.ptr, sizeof(int) * nT, cudaMemcpyDeviceToHost, st));
    }
    CU_CHECK(c, cudaStreamSynchronize(st));
    std::vector<float4>& par4 = c->w_par4;
    par4.resize(nT);
#pragma omp parallel for schedule(static) num_threads(host_threads(nT)) if (nT > c_host_grain)
    for (int k = 0; k < nT; k++)
    {
        const int a = c->touched[k];
        par4[k]     = make_float4(c->qA[a], c->qB[a], int_bits_as_float(c->typeA[a]), int_bits_as_float(c->typeB[a]));
    }
    int rc;
    if ((rc = to_device(c, c->d_par4, par4)))
    {
        return rc;
    }
    const size_t n_slots = 32 * (size_t)NT;
    CU_CHECK(c, c->d_trips.reserve(std::max<size_t>((size_t)NT * FEP_TRIP_WORDS, 1)));
    CU_CHECK(c, c->d_tgid.reserve(std::max(NT, 1)));
    CU_CHECK(c, c->d_tshift.reserve(std::max(NT, 1)));
    CU_CHECK(c, c->d_tfirst.reserve((size_t)NT + 1));
    CU_CHECK(c, c->d_kshift.reserve(std::max(NT, 1)));
    CU_CHECK(c, c->d_kgid.reserve(std::max(NT, 1)));
    CU_CHECK(c, c->d_orig.reserve(std::max<size_t>(n_slots, 1)));
    CU_CHECK(c, c->d_akeys.reserve(std::max(P + NT, 1)));
    CU_CHECK(c, c->d_akeys_out.reserve(std::max(P + NT, 1)));
    CU_CHECK(c, c->d_avals.reserve(std::max(P + NT, 1)));
    CU_CHECK(c, c->d_avals_out.reserve(std::max(P + NT, 1)));
    b.trips = c->d_trips.ptr, b.tgid = c->d_tgid.ptr, b.tshift = c->d_tshift.ptr, b.orig = c->d_orig.ptr;
    b.tfirst = c->d_tfirst.ptr, b.kshift = c->d_kshift.ptr, b.kgid = c->d_kgid.ptr;
    b.akeys = c->d_akeys.ptr, b.akeys_out = c->d_akeys_out.ptr, b.avals = c->d_avals.ptr, b.avals_out = c->d_avals_out.ptr;
    c->lb.b      = b;
    c->lb.d_excl = d_excl;
    c->lb.j0 = j0, c->lb.P = P, c->lb.NT = NT, c->lb.nT = nT, c->lb.ngrp = ngrp, c->lb.wide = wide ? 1 : 0;
    return FEPB200_OK;
}

/* trips per run (fep_types.h): the smallest power of two with which one wave of the pass kernel's resident
 * warps covers `trips`, at most FEP_MAX_RUN_TRIPS -- long runs mean few segments (few reductions, few owner
 * contributions), short runs mean parallelism for short lists */
static int choose_run_trips(const fepb200_ctx* c, long long trips)
{
    if (const char* e = std::getenv("FEPB200_RUN_TRIPS"))
    {
        const int v = std::atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8)
        {
            return v;
        }
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    const long long warps = (long long)sms * 8 * (FEP_FB_CTA / 32);
    int             r     = 1;
    while (r < FEP_MAX_RUN_TRIPS && trips > warps * r)
    {
        r *= 2;
    }
    return r;
}

/* host copies of what the device built: every atom's range in the atom-sorted buffer, the records of the light and
 * heavy atoms, and the atoms this context's pairs touch */
static int fetch_atom_lists(fepb200_ctx* c)
{
    if (c->host_lists_valid)
    {
        return FEPB200_OK;
    }
    const int nT = c->lb.nT;
    c->w_atom_ptr.assign((size_t)nT + 1, 0);
    c->light_atoms.resize(c->ka.n_light);
    c->heavy_atoms.resize(c->ka.n_heavy);
    if (c->lb.P > 0)
    {
        CU_CHECK(c, cudaMemcpyAsync(c->w_atom_ptr.data(), c->d_atom_ptr.ptr, sizeof(int) * ((size_t)nT + 1), cudaMemcpyDeviceToHost, c->stream));
        if (c->ka.n_light > 0)
        {
            CU_CHECK(c, cudaMemcpyAsync(c->light_atoms.data(), c->d_light.ptr, sizeof(int4) * (size_t)c->ka.n_light, cudaMemcpyDeviceToHost, c->stream));
        }
        if (c->ka.n_heavy > 0)
        {
            CU_CHECK(c, cudaMemcpyAsync(c->heavy_atoms.data(), c->d_heavy.ptr, sizeof(int4) * (size_t)c->ka.n_heavy, cudaMemcpyDeviceToHost, c->stream));
        }
        CU_CHECK(c, cudaStreamSynchronize(c->stream));
    }
    c->local_atoms.resize(c->light_atoms.size() + c->heavy_atoms.size());
    {
        /* merge of two ascending lists */
        size_t i = 0, j = 0, o = 0;
        while (i < c->light_atoms.size() || j < c->heavy_atoms.size())
        {
            if (j >= c->heavy_atoms.size() || (i < c->light_atoms.size() && c->light_atoms[i].x < c->heavy_atoms[j].x))
            {
                c->local_atoms[o++] = c->light_atoms[i++].x;
            }
            else
            {
                c->local_atoms[o++] = c->heavy_atoms[j++].x;
            }
        }
    }
    c->host_lists_valid = true;
    return FEPB200_OK;
}

/* Phase 3 of the list build for runs of `run_trips` trips, and everything that follows from it: segment slots,
 * atom ranges, reduction jobs, heavy atoms, the intermediate buffers.  set_list() calls it once; the peer
 * exchange calls it again when a rank's share of the trips wants shorter runs. */
static int build_segments(fepb200_ctx* c, int run_trips)
{
    cudaStream_t st = c->stream;
    const int    P = c->lb.P, NT = c->lb.NT, nT = c->lb.nT, ngrp = c->lb.ngrp;
    int          err = fep_list_build_slots(&c->lb.b, c->lb.d_excl, c->lb.j0, c->d_par4.ptr, c->ntype, P, NT, nT, ngrp, c->lb.wide,
                                            run_trips, st, &c->launches);
    if (err != 0)
    {
        return fail(c, FEPB200_ERR_CUDA, "list build (slots) failed: %s", cudaGetErrorString((cudaError_t)err));
    }
    /* the epilogue's atom records, light and heavy, compacted on the device */
    CU_CHECK(c, c->d_atom_rec.reserve(std::max(nT, 1)));
    CU_CHECK(c, c->d_light.reserve(std::max(nT, 1)));
    CU_CHECK(c, c->d_heavy.reserve(std::max(nT, 1)));
    CU_CHECK(c, c->d_rec_counts.reserve(2));
    err = fep_list_build_records(c->d_atom_ptr.ptr, P > 0 ? nT : 0, c->d_atom_rec.ptr, c->d_light.ptr, c->d_heavy.ptr,
                                 c->d_rec_counts.ptr, c->lb.b.tmp, c->lb.b.tmp_bytes, st, &c->launches);
    if (err != 0)
    {
        return fail(c, FEPB200_ERR_CUDA, "list build (records) failed: %s", cudaGetErrorString((cudaError_t)err));
    }
    /* back to the host: the per-key counts (for the reduction jobs), the number of contributions and of records */
    std::vector<int> key_ptr(FEP_NUM_SHIFT + 1 + ngrp + 1, 0);
    int              n_contrib = 0, counts[2] = { 0, 0 };
    CU_CHECK(c, cudaMemcpyAsync(key_ptr.data(), c->d_key_ptr.ptr, sizeof(int) * key_ptr.size(), cudaMemcpyDeviceToHost, st));
    CU_CHECK(c, cudaMemcpyAsync(&n_contrib, c->d_atom_ptr.ptr + nT, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_CHECK(c, cudaMemcpyAsync(counts, c->d_rec_counts.ptr, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_CHECK(c, cudaStreamSynchronize(st));
    const int NS = P > 0 ? n_contrib - P : 0;
    if (NS < 0 || NS > NT || key_ptr[FEP_NUM_SHIFT] != NS || key_ptr[FEP_NUM_SHIFT + 1 + ngrp] != NS)
    {
        return fail(c, FEPB200_ERR_CUDA, "list build: inconsistent segment counts (%d, %d, %d of %d trips)", NS,
                    key_ptr[FEP_NUM_SHIFT], key_ptr[FEP_NUM_SHIFT + 1 + ngrp], NT);
    }
    std::vector<RedJob> jobs;
    std::vector<int>    key_job_ptr;
    build_jobs(key_ptr.data(), key_ptr.data() + FEP_NUM_SHIFT + 1, ngrp, &jobs, &key_job_ptr, &c->ka.n_shift_jobs);
    int rc;
    if ((rc = to_device(c, c->d_red_jobs, jobs)) || (rc = to_device(c, c->d_key_job_ptr, key_job_ptr)))
    {
        return rc;
    }
    c->host_lists_valid = false;
    c->shard_local      = false;
    CU_CHECK(c, c->d_fsorted.reserve(std::max(P + NS, 1)));
    CU_CHECK(c, c->d_fshift_sorted.reserve(std::max(NS, 1)));
    CU_CHECK(c, c->d_ev2.reserve(std::max(NS, 1)));
    CU_CHECK(c, c->d_job_part.reserve(4 * std::max<size_t>(jobs.size(), 1)));
    CU_CHECK(c, cudaStreamSynchronize(st)); /* host vectors go out of scope */
    KernelArgs& k   = c->ka;
    c->result_needs_zero = true;
    c->own_on            = false;
    c->push_n            = 0; /* the receive blocks were sized (and are only ever partly written) for the previous list */
    c->zeroed_targets.clear();
    c->n_segs       = NS;
    k.n_segs        = NS;
    k.run_trips     = run_trips;
    k.n_red_jobs    = (int)jobs.size();
    k.fsorted       = c->d_fsorted.ptr;
    k.fshift_sorted = c->d_fshift_sorted.ptr;
    k.ev2           = c->d_ev2.ptr;
    k.job_part      = c->d_job_part.ptr;
    k.atom_ptr      = c->d_atom_ptr.ptr;
    k.heavy_atoms   = c->d_heavy.ptr;
    k.n_heavy       = counts[1];
    k.light_atoms   = c->d_light.ptr;
    k.n_light       = counts[0];
    k.red_jobs      = c->d_red_jobs.ptr;
    k.key_job_ptr   = c->d_key_job_ptr.ptr;
    return FEPB200_OK;
}

static bool create_copy_events(fepb200_ctx* c)
{
    for (int i = 0; i < c_copy_chunks; i++)
    {
        if (cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming) != cudaSuccess)
        {
            return false;
        }
    }
    return true;
}

/* =========================================================================================== */
extern "C" {

int fepb200_create(fepb200_ctx** out, int device_ordinal)
{
    if (!out)
    {
        return fail(nullptr, FEPB200_ERR_INVALID_ARGUMENT, "ctx pointer is NULL");
    }
    *out      = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    {
        return fail(nullptr, FEPB200_ERR_NO_DEVICE,
                    "no CUDA device available; libfepb200 has no CPU implementation of this path");
    }
    if (device_ordinal < 0 || device_ordinal >= count)
    {
        return fail(nullptr, FEPB200_ERR_INVALID_ARGUMENT, "device ordinal %d out of range [0,%d)", device_ordinal,
                    count);
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess)
    {
        return fail(nullptr, FEPB200_ERR_CUDA, "cudaGetDeviceProperties failed");
    }
    if (prop.major != 10)
    {
        return fail(nullptr, FEPB200_ERR_NO_DEVICE,
                    "device %d (%s) is sm_%d%d; this library is built for sm_100a only", device_ordinal, prop.name,
                    prop.major, prop.minor);
    }
    fepb200_ctx* c = new fepb200_ctx();
    c->lap_on      = std::getenv("FEPB200_TIMING") != nullptr;
    if (c->lap_on)
    {
        for (int i = 0; i < 4; i++)
        {
            cudaEventCreate(&c->lap_ev[i]);
        }
    }
    c->zc_out      = std::getenv("FEPB200_ZC_OUT") ? std::atoi(std::getenv("FEPB200_ZC_OUT")) != 0 : true;
    c->device      = device_ordinal;
    if (cudaSetDevice(device_ordinal) != cudaSuccess
        || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess
        || cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking) != cudaSuccess
        || cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming) != cudaSuccess
        || cudaEventCreateWithFlags(&c->join_ev, cudaEventDisableTiming) != cudaSuccess
        || cudaEventCreate(&c->ev_start) != cudaSuccess || cudaEventCreate(&c->ev_stop) != cudaSuccess
        || cudaEventCreateWithFlags(&c->ev_staged, cudaEventDisableTiming) != cudaSuccess
        || cudaEventCreateWithFlags(&c->ev_handoff, cudaEventDisableTiming) != cudaSuccess
        || !create_copy_events(c)
        || c->d_counter.reserve(1) != cudaSuccess || cudaMemset(c->d_counter.ptr, 0, sizeof(unsigned int)) != cudaSuccess)
    {
        const int rc = fail(nullptr, FEPB200_ERR_CUDA, "CUDA initialisation failed: %s",
                            cudaGetErrorString(cudaGetLastError()));
        delete c;
        return rc;
    }
    c->ka.done_counter = c->d_counter.ptr;
    c->own_stream      = c->stream;
    {
        /* best effort: without it kernels still trap, only the reason is not reported */
        void* hp = nullptr;
        void* dp = nullptr;
        if (cudaHostAlloc(&hp, FEP_FAULT_WORDS * sizeof(unsigned int), cudaHostAllocMapped) == cudaSuccess
            && cudaHostGetDevicePointer(&dp, hp, 0) == cudaSuccess)
        {
            c->h_fault = static_cast<unsigned int*>(hp);
            std::memset(c->h_fault, 0, FEP_FAULT_WORDS * sizeof(unsigned int));
            c->ka.fault = static_cast<unsigned int*>(dp);
        }
        else
        {
            if (hp)
            {
                cudaFreeHost(hp);
            }
            cudaGetLastError();
        }
    }
    char buf[256];
    snprintf(buf, sizeof(buf), "fepb200 0.1 sm_100a %s %d SMs", prop.name, prop.multiProcessorCount);
    c->description = buf;
    *out           = c;
    return FEPB200_OK;
}

int fepb200_destroy(fepb200_ctx* c)
{
    if (!c)
    {
        return FEPB200_OK;
    }
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->h_fault)
    {
        cudaFreeHost(c->h_fault);
        c->h_fault = nullptr;
    }
    if (c->lap_on && c->lap_n > 0)
    {
        const double n = (double)c->lap_n;
        std::fprintf(stderr,
                     "fepb200 compute() laps over %lld calls [us]: gather %.1f | H2D queued %.1f | kernels queued %.1f | "
                     "D2H queued %.1f | results arrived %.1f | scatter %.1f\n",
                     c->lap_n, c->lap_sum[0] / n, c->lap_sum[1] / n, c->lap_sum[2] / n, c->lap_sum[3] / n,
                     c->lap_sum[4] / n, c->lap_sum[5] / n);
        std::fprintf(stderr, "fepb200 compute() device intervals [us]: H2D %.1f | kernels %.1f | D2H %.1f\n",
                     c->lap_dev[0] / n, c->lap_dev[1] / n, c->lap_dev[2] / n);
    }
    c->d_typetab.release();
    c->d_pts.release();
    c->h_pts.release();
    c->d_touched.release();
    c->d_trips.release();
    c->d_orig.release();
    c->d_tgid.release();
    c->d_atom_ptr.release();
    c->d_key_job_ptr.release();
    c->d_ent4.release();
    c->d_red_jobs.release();
    c->d_par4.release();
    c->d_trace.release();
    c->d_heavy.release();
    c->d_light.release();
    c->d_slot_src.release();
    c->d_fshift_src.release();
    c->d_ev2_src.release();
    c->d_fsorted.release();
    c->d_fshift_sorted.release();
    c->d_ev2.release();
    c->d_cta_part.release();
    c->d_for_part.release();
    c->d_job_part.release();
    c->d_counter.release();
    c->d_raw_iinr.release();
    c->d_raw_gid.release();
    c->d_raw_shift.release();
    c->d_raw_jindex.release();
    c->d_raw_jjnr.release();
    c->d_raw_excl.release();
    c->d_mark.release();
    c->d_cscan.release();
    c->d_keys.release();
    c->d_keys_out.release();
    c->d_vals.release();
    c->d_vals_out.release();
    c->d_pj.release();
    c->d_pn.release();
    c->d_deg.release();
    c->d_gmark.release();
    c->d_gstart.release();
    c->d_th.release();
    c->d_tsc.release();
    c->d_akeys.release();
    c->d_akeys_out.release();
    c->d_avals.release();
    c->d_avals_out.release();
    c->d_tshift.release();
    c->d_key_ptr.release();
    c->d_cub_tmp.release();
    c->d_step_in.release();
    c->d_result.release();
    c->h_step_in.release();
    c->h_result.release();
    for (int i = 0; i < 4; i++)
    {
        if (c->ev_prof[i])
        {
            cudaEventDestroy(c->ev_prof[i]);
        }
    }
    for (cudaEvent_t e : c->ev_copy)
    {
        if (e)
        {
            cudaEventDestroy(e);
        }
    }
    cudaEventDestroy(c->ev_start);
    cudaEventDestroy(c->ev_stop);
    if (c->ev_handoff)
    {
        cudaEventDestroy(c->ev_handoff);
    }
    if (c->ev_staged)
    {
        cudaEventDestroy(c->ev_staged);
    }
    cudaStreamDestroy(c->own_stream);
    cudaStreamDestroy(c->side_stream);
    cudaEventDestroy(c->fork_ev);
    cudaEventDestroy(c->join_ev);
    delete c;
    return FEPB200_OK;
}

const char* fepb200_last_error(const fepb200_ctx* c)
{
    return c ? c->error.c_str() : g_create_error.c_str();
}

const char* fepb200_describe(const fepb200_ctx* c)
{
    return c ? c->description.c_str() : "fepb200 0.1 sm_100a (no context)";
}

int fepb200_set_stream(fepb200_ctx* c, void* stream)
{
    if (!c)
    {
        return FEPB200_ERR_INVALID_ARGUMENT;
    }
    cudaSetDevice(c->device);
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    c->staging_in_flight = false;
    c->stream            = stream ? static_cast<cudaStream_t>(stream) : c->own_stream;
    return FEPB200_OK;
}

int fepb200_set_params(fepb200_ctx* c, const fepb200_params* p)
{
    if (!c || !p)
    {
        return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "NULL argument");
    }
    /* nb_free_energy.cpp:1384-1386: only plain cut-off, reaction-field and Ewald-type electrostatics */
    const bool ewald = eel_is_ewald(p->eeltype);
    if (!(ewald || p->eeltype == FEPB200_EEL_CUT || eel_is_rf(p->eeltype)))
    {
        return fail(c, FEPB200_ERR_UNSUPPORTED, "eeltype %d is not supported by the perturbed-pair kernel",
                    p->eeltype);
    }
    if (p->lambdaPower != 1 && p->lambdaPower != 2)
    {
        return fail(c, FEPB200_ERR_UNSUPPORTED, "sc-power must be 1 or 2 (got %d)", p->lambdaPower);
    }
    if (p->softcoreType != FEPB200_SC_BEUTLER && p->softcoreType != FEPB200_SC_GAPSYS)
    {
        return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "unknown softcoreType %d", p->softcoreType);
    }
    cudaSetDevice(c->device);
    c->params     = *p;
    c->elec_ewald = ewald;
    /* nb_free_energy.cpp:1324-1363 */
    if (p->softcoreType == FEPB200_SC_BEUTLER)
    {
        c->softcore = (p->alphaCoulomb == 0.0f && p->alphaVdw == 0.0f) ? FEP_SC_NONE : FEP_SC_BEUTLER;
    }
    else
    {
        c->softcore = (p->gapsysScaleLinpointCoul == 0.0f && p->gapsysScaleLinpointVdW == 0.0f) ? FEP_SC_NONE
                                                                                                  : FEP_SC_GAPSYS;
    }
    KernelArgs& k = c->ka;
    k.epsfac      = p->epsfac;
    k.rcoulomb    = p->rcoulomb;
    k.rvdw        = p->rvdw;
    k.rvdw_switch = p->rvdw_switch;
    k.krf         = p->reactionFieldCoefficient;
    k.crf         = p->reactionFieldShift;
    k.sh_ewald    = p->sh_ewald;
    k.sh_lj_ewald = p->sh_lj_ewald;
    k.beta        = p->ewaldcoeff_q;
    k.beta2       = p->ewaldcoeff_q * p->ewaldcoeff_q;
    k.beta3       = k.beta2 * p->ewaldcoeff_q;
    k.lj_coeff_sq = p->ewaldcoeff_lj * p->ewaldcoeff_lj;
    k.lj_coeff6_div6 = k.lj_coeff_sq * k.lj_coeff_sq * k.lj_coeff_sq / 6.0f;
    k.disp_cpot      = p->dispersion_shift_cpot;
    k.rep_cpot       = p->repulsion_shift_cpot;
    const float rmax = std::max(p->rcoulomb, p->rvdw);
    k.rcut_max2      = rmax * rmax;
    {
        const double rc = p->rcoulomb, rv = p->rvdw;
        k.rcoulomb6 = (float)(rc * rc * rc * rc * rc * rc);
        k.rvdw6     = (float)(rv * rv * rv * rv * rv * rv);
    }
    k.alpha_c  = p->alphaCoulomb;
    k.alpha_v  = p->alphaVdw;
    k.gscale_c = p->gapsysScaleLinpointCoul;
    k.gscale_v = p->gapsysScaleLinpointVdW;
    k.gapsys_facel = p->epsfac;
    k.gapsys_rcoul = p->rcoulomb;
    k.vdw_ewald  = p->vdwtype == FEPB200_VDW_PME;
    k.pot_switch = p->vdw_modifier == FEPB200_MOD_POTSWITCH;
    k.rf_type    = !ewald;
    if (k.pot_switch)
    {
        /* nb_free_energy.cpp:361-370 */
        const double d = (double)p->rvdw - (double)p->rvdw_switch;
        k.sw_v3        = (float)(-10.0 / (d * d * d));
        k.sw_v4        = (float)(15.0 / (d * d * d * d));
        k.sw_v5        = (float)(-6.0 / (d * d * d * d * d));
        k.sw_f2        = (float)(-30.0 / (d * d * d));
        k.sw_f3        = (float)(60.0 / (d * d * d * d));
        k.sw_f4        = (float)(-30.0 / (d * d * d * d * d));
    }
    else
    {
        k.sw_v3 = k.sw_v4 = k.sw_v5 = k.sw_f2 = k.sw_f3 = k.sw_f4 = 0.0f;
    }
    c->have_params   = true;
    c->typetab_dirty = true;
    refresh_points(c);
    if (c->have_list)
    {
        const int rc = prepare_buffers(c);
        if (rc != FEPB200_OK)
        {
            return rc;
        }
        if (!c->pts.empty())
        {
            return upload_points(c);
        }
    }
    return FEPB200_OK;
}

int fepb200_set_nbfp(fepb200_ctx* c, int ntype, const float* nbfp, const float* nbfp_grid)
{
    if (!c || ntype <= 0 || !nbfp)
    {
        return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_set_nbfp: bad arguments");
    }
    cudaSetDevice(c->device);
    if (c->have_list && ntype != c->ntype)
    {
        c->have_list = false; /* types of the uploaded atoms refer to the old table */
    }
    c->ntype = ntype;
    c->nbfp.assign(nbfp, nbfp + 2 * (size_t)ntype * ntype);
    if (nbfp_grid)
    {
        c->nbfp_grid.assign(nbfp_grid, nbfp_grid + 2 * (size_t)ntype * ntype);
    }
    else
    {
        c->nbfp_grid.clear();
    }
    c->typetab_dirty = true;
    return FEPB200_OK;
}

int fepb200_set_atoms(fepb200_ctx* c, int natoms, const float* qA, const float* qB, const int* typeA,
                      const int* typeB)
{
    if (!c || natoms < 0 || (natoms > 0 && (!qA || !qB || !typeA || !typeB)))
    {
        return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_set_atoms: bad arguments");
    }
    if (c->ntype == 0)
    {
        return fail(c, FEPB200_ERR_STATE, "fepb200_set_nbfp() must precede fepb200_set_atoms()");
    }
    {
        int       bad  = -1;
        const int nthr = std::min(8, std::max(1, omp_get_max_threads()));
#pragma omp parallel for schedule(static) reduction(max : bad) num_threads(nthr) if (natoms > 65536)
        for (int a = 0; a < natoms; a++)
        {
            if (typeA[a] < 0 || typeA[a] >= c->ntype || typeB[a] < 0 || typeB[a] >= c->ntype)
            {
                bad = std::max(bad, a);
            }
        }
        if (bad >= 0)
        {
            return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "atom %d has a type outside [0,%d)", bad, c->ntype);
        }
        c->natoms = natoms;
        c->qA.resize(natoms);
        c->qB.resize(natoms);
        c->typeA.resize(natoms);
        c->typeB.resize(natoms);
        /* four arrays of natoms words: copied by four threads */
        if (natoms > 0)
        {
#pragma omp parallel sections num_threads(std::min(4, nthr)) if (natoms > 65536)
        {
#pragma omp section
            std::memcpy(c->qA.data(), qA, sizeof(float) * (size_t)natoms);
#pragma omp section
            std::memcpy(c->qB.data(), qB, sizeof(float) * (size_t)natoms);
#pragma omp section
            std::memcpy(c->typeA.data(), typeA, sizeof(int) * (size_t)natoms);
#pragma omp section
            std::memcpy(c->typeB.data(), typeB, sizeof(int) * (size_t)natoms);
        }
        }
    }
    c->have_list = false;
    return FEPB200_OK;
}

int fepb200_set_list(fepb200_ctx* c, int nri, const int* iinr, const int* gid, const int* shift, const int* jindex,
                     const int* jjnr, const int* excl_fep, int ngrp, int rank, int nranks)
{
    if (nri > 0 && (!iinr || !gid || !shift || !jindex))
    {
        return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_set_list: NULL list array");
    }
    fepb200_list_view v;
    v.nri = nri, v.iinr = iinr, v.gid = gid, v.shift = shift, v.jindex = jindex, v.jjnr = jjnr, v.excl_fep = excl_fep;
    return fepb200_set_lists(c, 1, &v, nullptr, 0, ngrp, rank, nranks);
}

int fepb200_set_lists(fepb200_ctx* c, int n_lists, const fepb200_list_view* lists, const int* atom_map, int n_map, int ngrp,
                      int rank, int nranks)
{
    if (!c || n_lists < 0 || (n_lists > 0 && !lists) || ngrp < 1 || nranks < 1 || rank < 0 || rank >= nranks || n_map < 0)
    {
        return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_set_list: bad arguments");
    }
    long long nri_ll = 0, nrj_total = 0;
    for (int l = 0; l < n_lists; l++)
    {
        const fepb200_list_view& v = lists[l];
        if (v.nri < 0 || (v.nri > 0 && (!v.iinr || !v.gid || !v.shift || !v.jindex)))
        {
            return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_set_list: NULL list array");
        }
        const long long np = v.nri > 0 ? v.jindex[v.nri] : 0;
        if (v.nri > 0 && (v.jindex[0] != 0 || np < 0 || (np > 0 && !v.jjnr)))
        {
            return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_set_list: jindex must start at 0");
        }
        if ((v.excl_fep != nullptr) != (lists[0].excl_fep != nullptr) && np > 0)
        {
            return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_set_lists: excl_fep must be given for all lists or for none");
        }
        nri_ll += v.nri;
        nrj_total += np;
    }
    if (c->natoms == 0 && nri_ll > 0)
    {
        return fail(c, FEPB200_ERR_STATE, "fepb200_set_atoms() must precede fepb200_set_list()");
    }
    if (nrj_total >= (1LL << 31) - 64 || nri_ll >= (1LL << 31) - 64)
    {
        return fail(c, FEPB200_ERR_UNSUPPORTED, "more than 2^31 pairs in one list");
    }
    const int nri = (int)nri_ll;
    cudaSetDevice(c->device);
    const bool  timing = std::getenv("FEPB200_TIMING") != nullptr;
    const auto  t_begin = std::chrono::steady_clock::now();
    auto        lap     = [&, last = t_begin](const char* what) mutable {
        if (timing)
        {
            const auto now = std::chrono::steady_clock::now();
            fprintf(stderr, "[fepb200_set_list] %-28s %8.3f ms\n", what,
                    std::chrono::duration<double, std::milli>(now - last).count());
            last = now;
        }
    };
    /* per i-entry checks on the host (O(nri)); the atom indices of the pairs are checked on the device */
    for (int l = 0; l < n_lists; l++)
    {
        const fepb200_list_view& v = lists[l];
        for (int n = 0; n < v.nri; n++)
        {
            if (v.jindex[n + 1] < v.jindex[n] || v.gid[n] < 0 || v.gid[n] >= ngrp || v.shift[n] < 0 || v.shift[n] >= FEP_NUM_SHIFT)
            {
                return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_set_list: i-entry %d of list %d is malformed", n, l);
            }
        }
    }

    lap("validate");
    /* this rank's contiguous range of i-entries of the concatenation, balanced by pair count */
    int       e0 = 0, e1 = nri;
    long long j0_ll = 0, j1_ll = nrj_total;
    if (nranks > 1)
    {
        std::vector<int>       first(nranks + 1, nri);
        std::vector<long long> first_pair(nranks + 1, nrj_total);
        first[0]               = 0;
        first_pair[0]          = 0;
        const long long target = (nrj_total + nranks - 1) / nranks;
        int             dest   = 0, n_glob = 0;
        long long       have   = 0, pairs_before = 0;
        for (int l = 0; l < n_lists; l++)
        {
            const fepb200_list_view& v = lists[l];
            for (int n = 0; n < v.nri; n++, n_glob++)
            {
                const long long nrj = v.jindex[n + 1] - v.jindex[n];
                if (dest + 1 < nranks && have > 0 && have + nrj - target > target - have)
                {
                    dest++;
                    first[dest]      = n_glob;
                    first_pair[dest] = pairs_before;
                    have             = 0;
                }
                have += nrj;
                pairs_before += nrj;
            }
        }
        e0 = first[rank], e1 = first[rank + 1];
        j0_ll = first_pair[rank], j1_ll = first_pair[rank + 1];
    }
    const int E  = e1 - e0;
    const int j0 = E > 0 ? (int)j0_ll : 0;
    const int P  = E > 0 ? (int)(j1_ll - j0_ll) : 0;

    /* the device layout is built on the GPU: kernels + scans + stable sorts (fep_list_build.cu) */
    int rc = build_list_device(c, n_lists, lists, atom_map, n_map, nri, nrj_total, ngrp, e0, E, j0, P);
    if (rc != FEPB200_OK)
    {
        return rc;
    }
    const int nT = c->lb.nT, H = c->lb.NT;
    lap("device build");
    KernelArgs& k = c->ka;
    c->px_on      = false; /* a new list has a new slot layout: the peer exchange must be set up again */
    k.px          = PeerExchange{};
    k.n_pairs     = P;
    k.n_entries   = E;
    k.n_trips     = H;
    k.n_touched   = nT;
    k.n_gid       = ngrp;
    k.par4        = c->d_par4.ptr;
    k.trips       = c->d_trips.ptr;
    if ((rc = build_segments(c, choose_run_trips(c, H))) != FEPB200_OK)
    {
        return rc;
    }
    if (nranks > 1)
    {
        /* a shard of a split list gathers and uploads the coordinates of ITS atoms only */
        if ((rc = fetch_atom_lists(c)) != FEPB200_OK)
        {
            return rc;
        }
        c->shard_local = (int)c->local_atoms.size() < nT;
    }
    lap("segments + buffers");

    fepb200_layout& l = c->layout;
    l.natoms          = c->natoms;
    l.ntouched        = nT;
    l.nri             = E;
    l.nrj             = P;
    l.nri_total       = nri;
    l.nrj_total       = nrj_total;
    l.nenergrp        = ngrp;
    l.nforeign        = (int)c->all_c.size();
    c->first_entry    = e0;
    c->n_trips        = H;
    c->have_list      = true;
    if ((rc = prepare_buffers(c)) != FEPB200_OK)
    {
        c->have_list = false;
        return rc;
    }
    if (!c->pts.empty())
    {
        return upload_points(c);
    }
    return FEPB200_OK;
}

int fepb200_get_list(const fepb200_ctx* cc, int* first_entry, int* iinr, int* gid, int* shift, int* jindex,
                     int* jjnr, int* excl_fep)
{
    fepb200_ctx* c = const_cast<fepb200_ctx*>(cc);
    if (!c || !c->have_list)
    {
        return fail(c, FEPB200_ERR_STATE, "no list has been set");
    }
    cudaSetDevice(c->device);
    if (first_entry)
    {
        *first_entry = c->first_entry;
    }
    const int E = c->layout.nri;
    const int P = (int)c->layout.nrj;
    /* read the list back FROM THE DEVICE and undo the compaction */
    if (iinr || gid || shift)
    {
        std::vector<int4> ent4(E);
        if (E > 0)
        {
            CU_CHECK(c, cudaMemcpy(ent4.data(), c->d_ent4.ptr, E * sizeof(int4), cudaMemcpyDeviceToHost));
        }
        for (int n = 0; n < E; n++)
        {
            if (iinr)
            {
                iinr[n] = c->touched[ent4[n].x];
            }
            if (shift)
            {
                shift[n] = ent4[n].y;
            }
            if (gid)
            {
                gid[n] = ent4[n].z;
            }
        }
    }
    if (jindex || jjnr || excl_fep)
    {
        /* jindex comes back from the raw copy the builder worked on; jjnr and excl_fep are RECONSTRUCTED from the
         * trip layout the kernels evaluate (slot -> original pair), and every pair's trip must agree with its
         * i-entry on (i atom, shift, gid): the read-back proves that the regrouping lost or changed nothing */
        const int           NT = c->n_trips;
        std::vector<int>          jraw((size_t)E + 1, 0), orig(32 * (size_t)NT), tgid(NT);
        std::vector<unsigned int> trips((size_t)NT * FEP_TRIP_WORDS);
        std::vector<int4>         ent4(E);
        if (E > 0)
        {
            CU_CHECK(c, cudaMemcpy(jraw.data(), c->d_raw_jindex.ptr + c->first_entry, ((size_t)E + 1) * sizeof(int),
                                   cudaMemcpyDeviceToHost));
            CU_CHECK(c, cudaMemcpy(ent4.data(), c->d_ent4.ptr, E * sizeof(int4), cudaMemcpyDeviceToHost));
        }
        if (NT > 0)
        {
            CU_CHECK(c, cudaMemcpy(trips.data(), c->d_trips.ptr, trips.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost));
            CU_CHECK(c, cudaMemcpy(orig.data(), c->d_orig.ptr, orig.size() * sizeof(int), cudaMemcpyDeviceToHost));
            CU_CHECK(c, cudaMemcpy(tgid.data(), c->d_tgid.ptr, NT * sizeof(int), cudaMemcpyDeviceToHost));
        }
        const int j0 = jraw[0];
        if (jindex)
        {
            for (int n = 0; n <= E; n++)
            {
                jindex[n] = jraw[n] - j0;
            }
        }
        std::vector<int>  entry_of(P);
        for (int n = 0; n < E; n++)
        {
            for (int k = jraw[n] - j0; k < jraw[n + 1] - j0; k++)
            {
                entry_of[k] = n;
            }
        }
        std::vector<char> seen(P, 0);
        long long         found = 0;
        for (size_t slot = 0; slot < orig.size(); slot++)
        {
            const unsigned int* tb  = trips.data() + (slot >> 5) * FEP_TRIP_WORDS;
            const unsigned int  cjx = tb[FEP_TW_CJX + (slot & 31)];
            if (cjx & FEP_SLOT_PADDING)
            {
                continue;
            }
            const int          sidx  = orig[slot];
            const unsigned int head  = tb[FEP_TH_OWNER];
            const bool         flip  = (head & FEP_TRIP_FLIPPED) != 0;
            const int          owner = (int)(head & (FEP_MAX_TOUCHED - 1)), other = (int)(cjx & (FEP_MAX_TOUCHED - 1));
            const int          sh_e  = (int)((head >> 24) & 63);
            const int  ci = flip ? other : owner, cj = flip ? owner : other;
            const int  sh = flip ? FEP_NUM_SHIFT - 1 - sh_e : sh_e;
            if (sidx < 0 || sidx >= P || seen[sidx])
            {
                return fail(c, FEPB200_ERR_STATE, "list layout is inconsistent: slot %zu names pair %d", slot, sidx);
            }
            const int4 en = ent4[entry_of[sidx]];
            if (en.x != ci || en.y != sh || en.z != tgid[slot >> 5])
            {
                return fail(c, FEPB200_ERR_STATE, "list layout is inconsistent: pair %d sits in a trip of another i-entry", sidx);
            }
            seen[sidx] = 1;
            found++;
            if (jjnr)
            {
                jjnr[sidx] = c->touched[cj];
            }
            if (excl_fep)
            {
                excl_fep[sidx] = (cjx & 0x80000000u) ? 0 : 1;
            }
        }
        if (found != P)
        {
            return fail(c, FEPB200_ERR_STATE, "list layout is inconsistent: %lld of %d pairs found", found, P);
        }
    }
    return FEPB200_OK;
}

int fepb200_touched_atoms(const fepb200_ctx* c, int* atoms)
{
    if (!c || !c->have_list || !atoms)
    {
        return fail(const_cast<fepb200_ctx*>(c), FEPB200_ERR_STATE, "no list has been set");
    }
    std::copy(c->touched.begin(), c->touched.end(), atoms);
    return FEPB200_OK;
}

int fepb200_result_layout(const fepb200_ctx* c, fepb200_layout* layout)
{
    if (!c || !layout || !c->have_list)
    {
        return fail(const_cast<fepb200_ctx*>(c), FEPB200_ERR_STATE, "no list has been set");
    }
    *layout = c->layout;
    return FEPB200_OK;
}

int fepb200_set_lambdas(fepb200_ctx* c, const float* lambda, int n_foreign, const float* all_c, const float* all_v)
{
    if (!c || !lambda || n_foreign < 0 || (n_foreign > 0 && (!all_c || !all_v)))
    {
        return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_set_lambdas: bad arguments");
    }
    if (n_foreign + 1 > FEP_MAX_POINTS)
    {
        return fail(c, FEPB200_ERR_UNSUPPORTED, "at most %d foreign lambda points", FEP_MAX_POINTS - 1);
    }
    cudaSetDevice(c->device);
    c->lam_c = lambda[FEPB200_LAMBDA_COUL];
    c->lam_v = lambda[FEPB200_LAMBDA_VDW];
    c->all_c.assign(all_c, all_c + n_foreign);
    c->all_v.assign(all_v, all_v + n_foreign);
    c->have_lambda = true;
    refresh_points(c);
    if (c->have_list)
    {
        /* sizes depend on L, the choice of foreign kernel on the lambda values */
        c->layout.nforeign = n_foreign;
        const int rc       = prepare_buffers(c);
        if (rc != FEPB200_OK)
        {
            return rc;
        }
        if (!c->pts.empty())
        {
            return upload_points(c);
        }
    }
    return FEPB200_OK;
}

/* ---- step ------------------------------------------------------------------------------- */
/* something other than the reduction kernel follows the epilogue on the stream: end the step here */
static void close_chain(fepb200_ctx* c)
{
    if (c->chain_open)
    {
        c->chain_open = false;
        cudaEventRecord(c->ev_stop, c->stream);
    }
}

static inline double wall_us()
{
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static inline void lap_us(fepb200_ctx* c, int i)
{
    if (c->lap_on)
    {
        const double t = wall_us();
        c->lap_sum[i] += t - c->lap_t0;
        c->lap_t0 = t;
    }
}

static int stage_head(fepb200_ctx* c, const float* shiftvec)
{
    DynHead* head = reinterpret_cast<DynHead*>(c->h_step_in.ptr);
    for (int s = 0; s < FEP_NUM_SHIFT; s++)
    {
        head->shiftvec[s] = make_float4(shiftvec[3 * s], shiftvec[3 * s + 1], shiftvec[3 * s + 2], 0.0f);
    }
    head->cur = c->cur;
    return FEPB200_OK;
}

int fepb200_upload_x(fepb200_ctx* c, const float* x, const float* shiftvec)
{
    if (!c || !x || !shiftvec)
    {
        return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_upload_x: NULL argument");
    }
    int rc = check_ready(c);
    if (rc != FEPB200_OK)
    {
        return rc;
    }
    cudaSetDevice(c->device);
    /* the previous H2D copy out of the pinned buffer must be done before it is overwritten (it is,
     * whenever a download or wait followed it: then no synchronisation is needed here) */
    if (c->staging_in_flight)
    {
        CU_CHECK(c, cudaStreamSynchronize(c->stream));
    }
    stage_head(c, shiftvec);
    float*     pos = reinterpret_cast<float*>(c->h_step_in.ptr + sizeof(DynHead));
    const int  nT  = c->layout.ntouched;
    const int* t   = c->touched.data();
    /* Pipelined: the touched coordinates are gathered into pinned memory chunk by chunk and each
     * chunk's H2D copy is queued at once, so the DMA of chunk k runs while the host gathers k+1. */
    if (c->shard_local && !c->px_on)
    {
        /* a shard of a split list: its kernels read the coordinates of the atoms of ITS pairs only -- gather those, and
         * copy the span of the compact array that holds them */
        const int* la = c->local_atoms.data();
        const int  nL = (int)c->local_atoms.size();
#pragma omp parallel for schedule(static) num_threads(host_threads(nT)) if (nL > c_host_grain)
        for (int i = 0; i < nL; i++)
        {
            const int    k  = la[i];
            const float* xa = x + 3 * (size_t)t[k];
            pos[3 * (size_t)k]     = xa[0];
            pos[3 * (size_t)k + 1] = xa[1];
            pos[3 * (size_t)k + 2] = xa[2];
        }
        lap_us(c, 0);
        CU_CHECK(c, cudaMemcpyAsync(c->d_step_in.ptr, c->h_step_in.ptr, sizeof(DynHead), cudaMemcpyHostToDevice, c->stream));
        if (nL > 0)
        {
            const size_t b0 = sizeof(DynHead) + 3 * sizeof(float) * (size_t)la[0];
            const size_t b1 = sizeof(DynHead) + 3 * sizeof(float) * ((size_t)la[nL - 1] + 1);
            CU_CHECK(c, cudaMemcpyAsync(c->d_step_in.ptr + b0, c->h_step_in.ptr + b0, b1 - b0, cudaMemcpyHostToDevice, c->stream));
        }
        lap_us(c, 1);
        c->staging_in_flight = true;
        c->staged_by_event   = false;
        return FEPB200_OK;
    }
    const int nchunks = nT > 32768 ? c_copy_chunks : 1;
    size_t    done    = 0; /* bytes of [DynHead | pos3] already queued */
    for (int ch = 0; ch < nchunks; ch++)
    {
        const int k0 = (int)((long long)nT * ch / nchunks), k1 = (int)((long long)nT * (ch + 1) / nchunks);
        /* one team size for the gather and the scatter of a context: libgomp tears surplus pool
         * threads down, and creates them again, whenever consecutive parallel regions ask for
         * different team sizes (measured: ~1 ms per step when a rank scatters only the atoms it owns) */
#pragma omp parallel for schedule(static) num_threads(host_threads(nT)) if (k1 - k0 > c_host_grain)
        for (int k = k0; k < k1; k++)
        {
            const float* xa = x + 3 * (size_t)t[k];
            pos[3 * (size_t)k]     = xa[0];
            pos[3 * (size_t)k + 1] = xa[1];
            pos[3 * (size_t)k + 2] = xa[2];
        }
        lap_us(c, 0);
        const size_t upto = sizeof(DynHead) + 3 * sizeof(float) * (size_t)k1;
        if (c->lap_on)
        {
            cudaEventRecord(c->lap_ev[0], c->stream);
        }
        CU_CHECK(c, cudaMemcpyAsync(c->d_step_in.ptr + done, c->h_step_in.ptr + done, upto - done, cudaMemcpyHostToDevice,
                                    c->stream));
        if (c->lap_on)
        {
            cudaEventRecord(c->lap_ev[1], c->stream);
        }
        done = upto;
    }
    lap_us(c, 1);
    c->staging_in_flight = true;
    c->staged_by_event   = false;
    return FEPB200_OK;
}

static int gather_x_device(fepb200_ctx* c, const float* d_x, int stride, const float* shiftvec)
{
    if (!c || !d_x || !shiftvec)
    {
        return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_gather_x_device: NULL argument");
    }
    int rc = check_ready(c);
    if (rc != FEPB200_OK)
    {
        return rc;
    }
    cudaSetDevice(c->device);
    if (c->staging_in_flight)
    {
        /* the previous copy out of the pinned head must be done before the head is overwritten */
        if (c->staged_by_event)
        {
            CU_CHECK(c, cudaEventSynchronize(c->ev_staged));
        }
        else
        {
            CU_CHECK(c, cudaStreamSynchronize(c->stream));
        }
    }
    stage_head(c, shiftvec);
    CU_CHECK(c, cudaMemcpyAsync(c->d_step_in.ptr, c->h_step_in.ptr, sizeof(DynHead), cudaMemcpyHostToDevice,
                                c->stream));
    CU_CHECK(c, cudaEventRecord(c->ev_staged, c->stream));
    c->staging_in_flight = true;
    c->staged_by_event   = true;
    const int err = fep_launch_gather_x(d_x, stride, c->d_touched.ptr,
                                        reinterpret_cast<float*>(c->d_step_in.ptr + sizeof(DynHead)),
                                        c->layout.ntouched, c->stream, &c->launches);
    if (err != 0)
    {
        return fail(c, FEPB200_ERR_CUDA, "gather kernel launch failed: %s", cudaGetErrorString((cudaError_t)err));
    }
    return FEPB200_OK;
}

int fepb200_gather_x_device(fepb200_ctx* c, const float* d_x, const float* shiftvec)
{
    return gather_x_device(c, d_x, 3, shiftvec);
}

int fepb200_gather_xq_device(fepb200_ctx* c, const float* d_xq, const float* shiftvec)
{
    return gather_x_device(c, d_xq, 4, shiftvec);
}

int fepb200_launch(fepb200_ctx* c, int flags, void* stream_v)
{
    if (!c)
    {
        return FEPB200_ERR_INVALID_ARGUMENT;
    }
    int rc = check_ready(c);
    if (rc != FEPB200_OK)
    {
        return rc;
    }
    cudaSetDevice(c->device);
    close_chain(c);
    cudaStream_t    stream = stream_v ? static_cast<cudaStream_t>(stream_v) : c->stream;
    const StepFlags sf     = step_flags(c, flags);
    if (stream != c->stream)
    {
        /* a launch stream of the caller's: the staging copies of upload_x / gather_x* (and the constants set earlier)
         * were queued on the context's stream -- the kernels must not start before them */
        CU_CHECK(c, cudaEventRecord(c->ev_handoff, c->stream));
        CU_CHECK(c, cudaStreamWaitEvent(stream, c->ev_handoff, 0));
    }
    CU_CHECK(c, cudaEventRecord(c->ev_start, stream));
    KernelArgs ka_step = c->ka;
    c->result_on_host  = false;
    if (c->res_target)
    {
        if (std::find(c->zeroed_targets.begin(), c->zeroed_targets.end(), (void*)c->res_target) == c->zeroed_targets.end())
        {
            CU_CHECK(c, cudaMemsetAsync(c->res_target, 0, c->res_f64_bytes + c->res_f32_bytes, stream));
            c->zeroed_targets.push_back(c->res_target);
        }
        ka_step.res_f64 = reinterpret_cast<double*>(c->res_target);
        ka_step.res_f32 = reinterpret_cast<float*>(c->res_target + c->res_f64_bytes);
    }
    else if (c->zc_next && c->h_result_dev)
    {
        /* fepb200_compute(): the epilogue writes the result block straight into the pinned host
         * buffer over PCIe (no separate D2H copy) */
        ka_step.res_f64   = reinterpret_cast<double*>(c->h_result_dev);
        ka_step.res_f32   = reinterpret_cast<float*>(c->h_result_dev + c->res_f64_bytes);
        c->result_on_host = true;
    }
    c->zc_next = false;
    if (c->push_n > 1)
    {
        /* push reduction: forces to the rank that owns the atom (the ranges of fepb200_reduce_scatter_peers), shift
         * forces and scalars to every rank; nothing of this step is written to a block of our own */
        const long long nT      = c->layout.ntouched;
        ka_step.push.nranks     = c->push_n;
        ka_step.push.per_rank   = (int)std::max(4LL, ((nT + c->push_n - 1) / c->push_n + 3) / 4 * 4);
        for (int r = 0; r < c->push_n; r++)
        {
            ka_step.push.f64[r] = reinterpret_cast<double*>(c->push_block[r]);
            ka_step.push.f32[r] = reinterpret_cast<float*>(c->push_block[r] + c->res_f64_bytes);
        }
        c->result_on_host = false;
    }
    if (c->px_on)
    {
        /* the exchange slot of this step on every rank; all ranks launch in lockstep, so they agree
         * on the sequence number.  Two alternating slots: a rank that runs ahead writes the other one
         * and cannot come back to this one before every rank has announced the next step. */
        const unsigned int seq  = ++c->x_seq;
        const size_t       so   = (size_t)(seq & 1u) * c->x_slot_bytes;
        PeerExchange&      px   = ka_step.px;
        px.nranks               = c->x_nranks;
        px.rank                 = c->x_rank;
        px.atom_begin           = c->x_atom_begin;
        px.atom_end             = c->x_atom_end;
        px.heavy_begin          = c->x_heavy_begin;
        px.heavy_end            = c->x_heavy_end;
        px.light_begin          = c->x_light_begin;
        px.light_end            = c->x_light_end;
        px.seq                  = seq;
        px.slot_src             = c->d_slot_src.ptr;
        px.fshift_src           = c->d_fshift_src.ptr;
        px.ev2_src              = c->d_ev2_src.ptr;
        for (int r = 0; r < c->x_nranks; r++)
        {
            unsigned char* b    = c->x_base[r] + so;
            px.fsorted[r]       = reinterpret_cast<float4*>(b + c->x_off_fsorted);
            px.fshift_sorted[r] = reinterpret_cast<float4*>(b + c->x_off_fshift);
            px.ev2[r]           = reinterpret_cast<float2*>(b + c->x_off_ev2);
            px.cta_part[r]      = reinterpret_cast<double*>(b + c->x_off_cta);
            px.for_part[r]      = reinterpret_cast<double*>(b + c->x_off_for);
            px.flags[r]         = reinterpret_cast<unsigned int*>(c->x_base[r] + 2 * c->x_slot_bytes);
        }
        /* the pair kernels of this rank write its own slot */
        ka_step.fsorted       = const_cast<float4*>(px.fsorted[c->x_rank]);
        ka_step.fshift_sorted = const_cast<float4*>(px.fshift_sorted[c->x_rank]);
        ka_step.ev2           = const_cast<float2*>(px.ev2[c->x_rank]);
        ka_step.cta_part      = const_cast<double*>(px.cta_part[c->x_rank]);
        ka_step.for_part      = const_cast<double*>(px.for_part[c->x_rank]);
    }
    const int err = fep_launch_step(&ka_step, c->softcore, c->elec_ewald, sf, stream, &c->launches,
                                    c->profiling ? c->ev_prof : nullptr, &c->cur, c->pts.data(), c->foreign_mode,
                                    c->side_stream, c->fork_ev, c->join_ev);
    c->profiled   = c->profiling;
    if (err != 0)
    {
        return fail(c, FEPB200_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString((cudaError_t)err));
    }
    /* multi-GPU with a published partial block: the reduction kernel follows on the same stream and
     * is chained behind the epilogue (nothing may be queued between them); it records ev_stop */
    c->chain_open = (c->res_target != nullptr || c->push_n > 1) && stream == c->stream && !c->profiling;
    if (!c->chain_open)
    {
        CU_CHECK(c, cudaEventRecord(c->ev_stop, stream));
    }
    if (stream != c->stream)
    {
        /* ... and every consumer (add_forces_device, export_scalars_device, download, the next staging copy) runs on
         * the context's stream: it must not start before the kernels are done */
        CU_CHECK(c, cudaStreamWaitEvent(c->stream, c->ev_stop, 0));
    }
    c->timed = true;
    return FEPB200_OK;
}

int fepb200_add_forces_device(fepb200_ctx* c, float* d_f, int flags)
{
    if (!c || !d_f)
    {
        return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_add_forces_device: NULL argument");
    }
    int rc = check_ready(c);
    if (rc != FEPB200_OK)
    {
        return rc;
    }
    if (c->result_on_host)
    {
        return fail(c, FEPB200_ERR_STATE,
                    "fepb200_add_forces_device: the last step was a fepb200_compute(), whose results live in host memory; "
                    "use fepb200_launch()");
    }
    cudaSetDevice(c->device);
    close_chain(c);
    /* the context's own result block: what fepb200_launch() fills (unless a partial block was set
     * for the peer reduction) and what fepb200_reduce_peers() leaves the sum over ranks in */
    const float* r32 = c->ka.res_f32;
    const int    k0  = c->px_on ? c->x_atom_begin : (c->own_on ? c->own_begin : 0);
    const int    k1  = c->px_on ? c->x_atom_end : (c->own_on ? c->own_end : c->layout.ntouched);
    const int    err = fep_launch_add_forces(r32, c->d_touched.ptr, d_f, k0, k1,
                                             (flags & FEPB200_CLEAR_OUTPUTS) != 0 ? FEP_ADD_OVERWRITE
                                             : ((flags & FEPB200_ATOMIC_OUTPUTS) != 0 ? FEP_ADD_ATOMIC : FEP_ADD_PLAIN),
                                             c->stream, &c->launches);
    if (err != 0)
    {
        return fail(c, FEPB200_ERR_CUDA, "force scatter launch failed: %s", cudaGetErrorString((cudaError_t)err));
    }
    return FEPB200_OK;
}

int fepb200_export_scalars_device(fepb200_ctx* c, int flags, float* eLJ, float* eElec, float* dvdlLJ, float* dvdlElec,
                                  float* eLJForeign, float* eElecForeign, float* dvdlLJForeign, float* dvdlElecForeign,
                                  float* fShift)
{
    if (!c)
    {
        return FEPB200_ERR_INVALID_ARGUMENT;
    }
    int rc = check_ready(c);
    if (rc != FEPB200_OK)
    {
        return rc;
    }
    if (c->result_on_host)
    {
        return fail(c, FEPB200_ERR_STATE,
                    "fepb200_export_scalars_device: the last step was a fepb200_compute(), whose results live in host memory; "
                    "use fepb200_launch()");
    }
    cudaSetDevice(c->device);
    close_chain(c);
    const StepFlags       sf = step_flags(c, flags);
    const fepb200_layout& l  = c->layout;
    ExportLayout          lay;
    lay.ngrp             = l.nenergrp;
    lay.nforeign         = l.nforeign;
    lay.energy           = sf.energy ? 1 : 0;
    lay.foreign          = sf.foreign ? 1 : 0;
    lay.shift            = (sf.force && sf.shift) ? 1 : 0;
    lay.atomic           = (flags & FEPB200_ATOMIC_OUTPUTS) != 0 ? 1 : 0;
    lay.off_vc           = (int)l.off_vc;
    lay.off_vv           = (int)l.off_vv;
    lay.off_dvdl         = (int)l.off_dvdl;
    lay.off_foreign_e    = (int)l.off_foreign_e;
    lay.off_foreign_dvdl = (int)l.off_foreign_dvdl;
    ExportTargets t{ eLJ, eElec, dvdlLJ, dvdlElec, eLJForeign, eElecForeign, dvdlLJForeign, dvdlElecForeign, fShift };
    (void)eElecForeign; /* the whole energy of a point goes through eLJForeign, see the kernel */
    const int err = fep_launch_export_scalars(c->ka.res_f64, c->ka.res_f32 + l.off_fshift, &lay, &t, c->stream, &c->launches);
    if (err != 0)
    {
        return fail(c, FEPB200_ERR_CUDA, "scalar export launch failed: %s", cudaGetErrorString((cudaError_t)err));
    }
    return FEPB200_OK;
}

int fepb200_wait(fepb200_ctx* c)
{
    if (!c)
    {
        return FEPB200_ERR_INVALID_ARGUMENT;
    }
    cudaSetDevice(c->device);
    close_chain(c);
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    c->staging_in_flight = false;
    return FEPB200_OK;
}

int fepb200_result_device_ptrs(const fepb200_ctx* c, void** d_f32, void** d_f64)
{
    if (!c || !c->have_list)
    {
        return fail(const_cast<fepb200_ctx*>(c), FEPB200_ERR_STATE, "no list has been set");
    }
    if (d_f32)
    {
        *d_f32 = c->ka.res_f32;
    }
    if (d_f64)
    {
        *d_f64 = c->ka.res_f64;
    }
    return FEPB200_OK;
}

size_t fepb200_result_block_bytes(const fepb200_ctx* c)
{
    return (c && c->have_list) ? c->res_f64_bytes + c->res_f32_bytes : 0;
}

int fepb200_set_partial_result_block(fepb200_ctx* c, void* d_block)
{
    if (!c || !c->have_list)
    {
        return fail(c, FEPB200_ERR_STATE, "no list has been set");
    }
    if ((reinterpret_cast<size_t>(d_block) & 15) != 0)
    {
        return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "result block must be 16-byte aligned");
    }
    c->res_target = static_cast<unsigned char*>(d_block);
    return FEPB200_OK;
}

int fepb200_set_push_targets(fepb200_ctx* c, int nranks, void* const* d_peer_blocks)
{
    if (!c || !c->have_list)
    {
        return fail(c, FEPB200_ERR_STATE, "no list has been set");
    }
    if (nranks <= 1 || !d_peer_blocks)
    {
        c->push_n = 0;
        return FEPB200_OK;
    }
    if (nranks > FEP_XMAX)
    {
        return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_set_push_targets: at most %d ranks", FEP_XMAX);
    }
    if (c->px_on)
    {
        return fail(c, FEPB200_ERR_STATE, "fepb200_set_push_targets: the fused peer exchange is on (it is the reduction)");
    }
    for (int r = 0; r < nranks; r++)
    {
        if (!d_peer_blocks[r] || (reinterpret_cast<size_t>(d_peer_blocks[r]) & 15) != 0)
        {
            return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_set_push_targets: block %d is NULL or not 16-byte aligned", r);
        }
    }
    for (int r = 0; r < nranks; r++)
    {
        c->push_block[r] = static_cast<unsigned char*>(d_peer_blocks[r]);
    }
    c->push_n = nranks;
    return FEPB200_OK;
}

int fepb200_publish_result(fepb200_ctx* c, void* d_block)
{
    if (!c || !c->have_list || !d_block)
    {
        return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_publish_result: bad arguments");
    }
    cudaSetDevice(c->device);
    close_chain(c);
    /* the device result block is contiguous [f64 | f32] */
    CU_CHECK(c, cudaMemcpyAsync(d_block, c->d_result.ptr, c->res_f64_bytes + c->res_f32_bytes, cudaMemcpyDeviceToDevice,
                                c->stream));
    return FEPB200_OK;
}

int fepb200_reduce_peers(fepb200_ctx* c, int nranks, void* const* d_peer_blocks, void* const* d_peer_flags, int rank,
                         unsigned int seq)
{
    if (!c || !c->have_list || !d_peer_blocks || nranks < 1 || nranks > FEP_MAX_PEERS || rank < 0 || rank >= nranks)
    {
        return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_reduce_peers: bad arguments (at most %d ranks)",
                    FEP_MAX_PEERS);
    }
    cudaSetDevice(c->device);
    c->own_on = false; /* every rank receives everything */
    PeerPtrs pp{}, ff{};
    for (int r = 0; r < nranks; r++)
    {
        pp.p[r] = d_peer_blocks[r];
        ff.p[r] = d_peer_flags ? d_peer_flags[r] : nullptr;
    }
    const int err = fep_launch_peer_reduce(&pp, d_peer_flags ? &ff : nullptr, rank, seq, nranks, c->ka.res_f64,
                                           (int)c->layout.f64_words, c->res_f64_bytes, c->ka.res_f32,
                                           c->layout.f32_words, c->stream, &c->launches, c->chain_open ? 1 : 0, c->ka.fault);
    if (c->chain_open)
    {
        c->chain_open = false;
        cudaEventRecord(c->ev_stop, c->stream);
    }
    if (err != 0)
    {
        return fail(c, FEPB200_ERR_CUDA, "peer reduce launch failed: %s", cudaGetErrorString((cudaError_t)err));
    }
    return FEPB200_OK;
}

int fepb200_reduce_scatter_peers(fepb200_ctx* c, int nranks, void* const* d_peer_blocks, void* const* d_peer_flags, int rank,
                                 unsigned int seq)
{
    if (!c || !c->have_list || !d_peer_blocks || nranks < 1 || nranks > FEP_MAX_PEERS || rank < 0 || rank >= nranks)
    {
        return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_reduce_scatter_peers: bad arguments (at most %d ranks)",
                    FEP_MAX_PEERS);
    }
    if (c->px_on)
    {
        return fail(c, FEPB200_ERR_STATE, "fepb200_reduce_scatter_peers: the fused peer exchange is on (it is the reduction)");
    }
    cudaSetDevice(c->device);
    PeerPtrs pp{}, ff{};
    for (int r = 0; r < nranks; r++)
    {
        pp.p[r] = d_peer_blocks[r];
        ff.p[r] = d_peer_flags ? d_peer_flags[r] : nullptr;
    }
    /* equal atom ranges, starting on multiples of four atoms (so that a range of 3-word forces starts on a 16-byte
     * boundary): the compact numbering is the one of the full list on every rank */
    const long long nT  = c->layout.ntouched;
    const long long per = ((nT + nranks - 1) / nranks + 3) / 4 * 4;
    const int       a0  = (int)std::min(nT, per * rank), a1 = (int)std::min(nT, per * (rank + 1));
    bool chained = c->chain_open;
    if (!c->own_on || c->own_begin != a0 || c->own_end != a1)
    {
        /* forces of atoms other ranks own are never written here: keep them zero (something between the epilogue
         * and the reduction: this one launch is not chained) */
        CU_CHECK(c, cudaMemsetAsync(c->d_result.ptr, 0, c->res_f64_bytes + c->res_f32_bytes, c->stream));
        chained      = false;
        c->own_on    = true;
        c->own_begin = a0;
        c->own_end   = a1;
    }
    const int err = fep_launch_peer_reduce_scatter(&pp, d_peer_flags ? &ff : nullptr, rank, seq, nranks, c->ka.res_f64,
                                                   (int)c->layout.f64_words, c->res_f64_bytes, c->ka.res_f32, 3LL * a0, 3LL * a1,
                                                   c->layout.off_fshift, c->stream, &c->launches, chained ? 1 : 0, c->ka.fault);
    if (c->chain_open)
    {
        c->chain_open = false;
        cudaEventRecord(c->ev_stop, c->stream);
    }
    if (err != 0)
    {
        return fail(c, FEPB200_ERR_CUDA, "peer reduce-scatter launch failed: %s", cudaGetErrorString((cudaError_t)err));
    }
    return FEPB200_OK;
}

/* back to the plain single-GPU path: launch geometry of the whole list, results in the context's own buffers */
static int peer_exchange_off(fepb200_ctx* c)
{
    c->px_on    = false;
    c->x_nranks = 1;
    c->x_rank   = 0;
    int rc      = FEPB200_OK;
    if (choose_run_trips(c, c->ka.n_trips) != c->ka.run_trips && (rc = build_segments(c, choose_run_trips(c, c->ka.n_trips))) != FEPB200_OK)
    {
        return rc;
    }
    rc = prepare_buffers(c);
    if (rc != FEPB200_OK)
    {
        return rc;
    }
    if (!c->pts.empty())
    {
        return upload_points(c);
    }
    return FEPB200_OK;
}

size_t fepb200_exchange_bytes(const fepb200_ctx* c, int nranks)
{
    if (!c || !c->have_list || nranks < 1 || nranks > FEP_XMAX)
    {
        return 0;
    }
    /* upper bound that does not depend on the launch geometry or the run length: at most one partial sum per
     * trip of a rank's share, room for 32 lambda points (or the current number if larger), one segment per trip */
    auto            up      = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t    P       = (size_t)c->ka.n_pairs, H = (size_t)c->ka.n_trips;
    const long long tpr     = ((long long)H + nranks - 1) / nranks + FEP_MAX_RUN_TRIPS;
    const size_t    ctas    = (size_t)(4 * (tpr + 1)); /* the force-only pass writes four sums per warp */
    const size_t    np      = (size_t)std::max(c->layout.nforeign + 1, 32);
    const size_t    slot    = up((P + H) * sizeof(float4)) + up(H * sizeof(float4)) + up(H * sizeof(float2))
                        + up(4 * ctas * sizeof(double)) + up(3 * np * ctas * sizeof(double));
    return 2 * slot + 256;
}

int fepb200_set_peer_exchange(fepb200_ctx* c, int nranks, int rank, void* const* d_peer_bufs, size_t bytes)
{
    if (!c || !c->have_list)
    {
        return fail(c, FEPB200_ERR_STATE, "fepb200_set_peer_exchange: no list has been set");
    }
    if (nranks < 1 || nranks > FEP_XMAX || rank < 0 || rank >= nranks || (nranks > 1 && !d_peer_bufs))
    {
        return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_set_peer_exchange: bad arguments (at most %d ranks)", FEP_XMAX);
    }
    if (c->layout.nri != c->layout.nri_total)
    {
        return fail(c, FEPB200_ERR_STATE,
                    "fepb200_set_peer_exchange: every rank must hold the full list (fepb200_set_list with rank 0 of 1)");
    }
    cudaSetDevice(c->device);
    close_chain(c);
    KernelArgs& k = c->ka;
    const int   P = k.n_pairs, H = k.n_trips, nT = k.n_touched;
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    if (nranks == 1)
    {
        return peer_exchange_off(c);
    }
    for (int r = 0; r < nranks; r++)
    {
        if (!d_peer_bufs[r] || (reinterpret_cast<size_t>(d_peer_bufs[r]) & 255) != 0)
        {
            return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_set_peer_exchange: buffer %d is NULL or not 256-byte aligned", r);
        }
    }
    c->px_on                = true;
    c->x_nranks             = nranks;
    c->x_rank               = rank;
    {
        for (int r = 0; r < nranks; r++)
        {
            c->x_base[r] = static_cast<unsigned char*>(d_peer_bufs[r]);
        }
        c->x_bytes = bytes;
        /* pairs: equal shares of the trips, in whole runs; a rank's share of the list may want shorter runs than
         * the whole list did (parallelism), which changes the segments and with them every slot */
        {
            const int want = choose_run_trips(c, ((long long)H + nranks - 1) / nranks);
            if (want != k.run_trips)
            {
                const int rc_seg = build_segments(c, want);
                if (rc_seg != FEPB200_OK)
                {
                    c->px_on = false;
                    return rc_seg;
                }
            }
        }
        const long long R       = k.run_trips;
        const long long tpr     = (((long long)H + nranks - 1) / nranks + R - 1) / R * R;
        c->x_range_trips        = (int)tpr;
        c->x_trip_begin         = (int)std::min<long long>((long long)rank * tpr, H);
        c->x_trip_end           = (int)std::min<long long>((long long)(rank + 1) * tpr, H);
        /* atoms: contiguous ranges with equal shares of the force contributions */
        {
            const int rc_lists = fetch_atom_lists(c);
            if (rc_lists != FEPB200_OK)
            {
                c->px_on = false;
                return rc_lists;
            }
        }
        const std::vector<int>& atom_ptr = c->w_atom_ptr;
        /* cost of an atom = its contributions + a fixed share for the lanes that serve it (a range
         * of many light atoms needs more blocks than a range of few heavy ones with the same volume) */
        const long long per_atom = 8;
        const long long total    = (long long)P + c->n_segs + per_atom * nT;
        int             a_of[FEP_XMAX + 1];
        a_of[0]      = 0;
        a_of[nranks] = nT;
        {
            int a = 0;
            for (int r = 1; r < nranks; r++)
            {
                const long long target = total * r / nranks;
                while (a < nT && atom_ptr[a] + per_atom * a < target)
                {
                    a++;
                }
                a_of[r] = a;
            }
        }
        c->x_atom_begin  = a_of[rank];
        c->x_atom_end    = a_of[rank + 1];
        auto first_at = [](const std::vector<int4>& v, int atom) {
            return (int)(std::lower_bound(v.begin(), v.end(), atom, [](const int4& r, int a) { return r.x < a; }) - v.begin());
        };
        c->x_heavy_begin = first_at(c->heavy_atoms, a_of[rank]);
        c->x_heavy_end   = first_at(c->heavy_atoms, a_of[rank + 1]);
        c->x_light_begin = first_at(c->light_atoms, a_of[rank]);
        c->x_light_end   = first_at(c->light_atoms, a_of[rank + 1]);
        /* producer rank of every element of the sorted arrays */
        CU_CHECK(c, c->d_slot_src.reserve(std::max<size_t>((size_t)P + H, 1)));
        CU_CHECK(c, c->d_fshift_src.reserve(std::max(H, 1)));
        CU_CHECK(c, c->d_ev2_src.reserve(std::max(H, 1)));
        const int err = fep_launch_source_tables(c->d_trips.ptr, H, (int)tpr, c->d_slot_src.ptr, c->d_fshift_src.ptr,
                                                 c->d_ev2_src.ptr, c->stream, &c->launches);
        if (err != 0)
        {
            c->px_on = false;
            return fail(c, FEPB200_ERR_CUDA, "source tables failed: %s", cudaGetErrorString((cudaError_t)err));
        }
    }
    c->x_seq = 0;
    int rc   = prepare_buffers(c);
    if (rc != FEPB200_OK)
    {
        const std::string why = c->error;
        peer_exchange_off(c);
        c->error = why;
        return rc;
    }
    /* forces of atoms other ranks own are never written here: keep them zero */
    CU_CHECK(c, cudaMemsetAsync(c->d_result.ptr, 0, c->res_f64_bytes + c->res_f32_bytes, c->stream));
    std::memset(c->h_result.ptr, 0, c->res_f64_bytes + c->res_f32_bytes);
    if (!c->pts.empty() && (rc = upload_points(c)) != FEPB200_OK)
    {
        return rc;
    }
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    return FEPB200_OK;
}

int fepb200_epilogue_trace(fepb200_ctx* c, int enable, unsigned long long* stamps, int max_blocks)
{
    if (!c || !c->have_list)
    {
        return fail(c, FEPB200_ERR_STATE, "no list has been set");
    }
    cudaSetDevice(c->device);
    int n = 0;
    if (stamps && max_blocks > 0 && c->ka.trace)
    {
        CU_CHECK(c, cudaStreamSynchronize(c->stream));
        n = std::min(max_blocks, FEP_TRACE_BLOCKS);
        CU_CHECK(c, cudaMemcpy(stamps, c->d_trace.ptr, sizeof(unsigned long long) * 4 * (size_t)n, cudaMemcpyDeviceToHost));
    }
    if (enable && !c->ka.trace)
    {
        CU_CHECK(c, c->d_trace.reserve(4 * (size_t)FEP_TRACE_BLOCKS));
        CU_CHECK(c, cudaMemset(c->d_trace.ptr, 0, sizeof(unsigned long long) * 4 * FEP_TRACE_BLOCKS));
        c->ka.trace = c->d_trace.ptr;
    }
    else if (!enable)
    {
        c->ka.trace = nullptr;
    }
    return n;
}

int fepb200_peer_ranges(const fepb200_ctx* c, int* pair_begin, int* pair_end, int* atom_begin, int* atom_end)
{
    if (!c || !c->have_list)
    {
        return fail(const_cast<fepb200_ctx*>(c), FEPB200_ERR_STATE, "no list has been set");
    }
    /* the range of trips (32 pair slots each) this context evaluates */
    if (pair_begin) *pair_begin = c->px_on ? c->x_trip_begin : 0;
    if (pair_end) *pair_end = c->px_on ? c->x_trip_end : c->ka.n_trips;
    if (atom_begin) *atom_begin = c->px_on ? c->x_atom_begin : (c->own_on ? c->own_begin : 0);
    if (atom_end) *atom_end = c->px_on ? c->x_atom_end : (c->own_on ? c->own_end : c->ka.n_touched);
    return FEPB200_OK;
}

int fepb200_download(fepb200_ctx* c, int flags, float* f, float* fshift, double* Vc, double* Vv, double* dvdl,
                     double* foreign_energy, double* foreign_dvdl)
{
    if (!c)
    {
        return FEPB200_ERR_INVALID_ARGUMENT;
    }
    int rc = check_ready(c);
    if (rc != FEPB200_OK)
    {
        return rc;
    }
    cudaSetDevice(c->device);
    close_chain(c);
    const StepFlags       sf    = step_flags(c, flags);
    const fepb200_layout& l     = c->layout;
    const bool            clear = (flags & FEPB200_CLEAR_OUTPUTS) != 0;
    if ((sf.force && !f) || (sf.shift && !fshift) || (sf.energy && (!Vc || !Vv)) || !dvdl
        || (sf.foreign && (!foreign_energy || !foreign_dvdl)))
    {
        return fail(c, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_download: an output requested by flags is NULL");
    }
    /* Pipelined: the fp64 block and the compact forces come back in chunks, each followed by an
     * event; the host scatter-adds chunk k while the DMA of chunk k+1 runs. */
    const int    nT      = l.ntouched;
    /* peer exchange: this rank holds the forces of the atoms it owns (and all scalars) */
    const bool   owned   = c->px_on || c->own_on;
    const int    ka0     = c->px_on ? c->x_atom_begin : (c->own_on ? c->own_begin : 0);
    const int    ka1     = c->px_on ? c->x_atom_end : (c->own_on ? c->own_end : nT);
    const int    nA      = ka1 - ka0;
    const int    nchunks = (sf.force && nA > 32768) ? c_copy_chunks : 1;
    const size_t f32_off = c->res_f64_bytes;
    if (owned && !c->result_on_host)
    {
        /* fp64 block and shift forces, then the owned force range in chunks */
        CU_CHECK(c, cudaMemcpyAsync(c->h_result.ptr, c->d_result.ptr, f32_off, cudaMemcpyDeviceToHost, c->stream));
        const size_t so = f32_off + sizeof(float) * (size_t)l.off_fshift;
        CU_CHECK(c, cudaMemcpyAsync(c->h_result.ptr + so, c->d_result.ptr + so, sizeof(float) * 3 * FEP_NUM_SHIFT,
                                    cudaMemcpyDeviceToHost, c->stream));
    }
    {
        size_t done = 0;
        for (int ch = 0; ch < nchunks; ch++)
        {
            size_t upto = f32_off;
            if (sf.force)
            {
                const int k1 = ka0 + (int)((long long)nA * (ch + 1) / nchunks);
                upto         = (ch == nchunks - 1 && !owned) ? f32_off + c->res_f32_bytes
                                                                : f32_off + sizeof(float) * 3 * (size_t)k1;
            }
            if (owned && ch == 0)
            {
                done = f32_off + sizeof(float) * 3 * (size_t)ka0;
            }
            if (!c->result_on_host && upto > done)
            {
                CU_CHECK(c, cudaMemcpyAsync(c->h_result.ptr + done, c->d_result.ptr + done, upto - done,
                                            cudaMemcpyDeviceToHost, c->stream));
            }
            CU_CHECK(c, cudaEventRecord(c->ev_copy[ch], c->stream));
            done = upto;
        }
    }
    if (c->lap_on)
    {
        cudaEventRecord(c->lap_ev[3], c->stream);
    }
    lap_us(c, 3);
    const double* r64 = reinterpret_cast<const double*>(c->h_result.ptr);
    const float*  r32 = reinterpret_cast<const float*>(c->h_result.ptr + c->res_f64_bytes);
    if (sf.force)
    {
        /* scatter of the compact forces into the caller's rvec array; only atoms that occur in
         * the list are written (with FEPB200_CLEAR_OUTPUTS: overwritten, all others untouched) */
        const int* t = c->touched.data();
        for (int ch = 0; ch < nchunks; ch++)
        {
            const int k0 = ka0 + (int)((long long)nA * ch / nchunks), k1 = ka0 + (int)((long long)nA * (ch + 1) / nchunks);
            CU_CHECK(c, cudaEventSynchronize(c->ev_copy[ch]));
            lap_us(c, 4);
            /* same team size as the gather of fepb200_upload_x() (see there) */
#pragma omp parallel for schedule(static) num_threads(host_threads(nT)) if (k1 - k0 > c_host_grain)
            for (int k = k0; k < k1; k++)
            {
                float* fa = f + 3 * (size_t)t[k];
                if (clear)
                {
                    fa[0] = r32[3 * k];
                    fa[1] = r32[3 * k + 1];
                    fa[2] = r32[3 * k + 2];
                }
                else
                {
                    fa[0] += r32[3 * k];
                    fa[1] += r32[3 * k + 1];
                    fa[2] += r32[3 * k + 2];
                }
            }
        }
        if (sf.shift)
        {
            for (int i = 0; i < 3 * FEP_NUM_SHIFT; i++)
            {
                fshift[i] = (clear ? 0.0f : fshift[i]) + r32[l.off_fshift + i];
            }
        }
    }
    CU_CHECK(c, cudaEventSynchronize(c->ev_copy[nchunks - 1]));
    lap_us(c, 5);
    c->staging_in_flight = false; /* everything queued before the last D2H copy has completed */
    if (sf.energy)
    {
        for (int g = 0; g < l.nenergrp; g++)
        {
            Vc[g] = (clear ? 0.0 : Vc[g]) + r64[l.off_vc + g];
            Vv[g] = (clear ? 0.0 : Vv[g]) + r64[l.off_vv + g];
        }
    }
    dvdl[0] = (clear ? 0.0 : dvdl[0]) + r64[l.off_dvdl];
    dvdl[1] = (clear ? 0.0 : dvdl[1]) + r64[l.off_dvdl + 1];
    if (sf.foreign)
    {
        /* the reference stores (not accumulates) these per lambda point, freeenergydispatch.cpp:298-305,
         * but several localities add up in ForeignLambdaTerms::accumulate: keep += semantics */
        for (int i = 0; i <= l.nforeign; i++)
        {
            foreign_energy[i] = (clear ? 0.0 : foreign_energy[i]) + r64[l.off_foreign_e + i];
            foreign_dvdl[2 * i] = (clear ? 0.0 : foreign_dvdl[2 * i]) + r64[l.off_foreign_dvdl + 2 * i];
            foreign_dvdl[2 * i + 1] =
                    (clear ? 0.0 : foreign_dvdl[2 * i + 1]) + r64[l.off_foreign_dvdl + 2 * i + 1];
        }
    }
    return FEPB200_OK;
}

int fepb200_compute(fepb200_ctx* c, const float* x, const float* shiftvec, int flags, float* f, float* fshift,
                    double* Vc, double* Vv, double* dvdl, double* foreign_energy, double* foreign_dvdl)
{
    if (c && c->lap_on)
    {
        c->lap_t0 = wall_us();
        c->lap_n++;
    }
    int rc = fepb200_upload_x(c, x, shiftvec);
    if (rc != FEPB200_OK)
    {
        return rc;
    }
    c->zc_next = c->zc_out;
    rc         = fepb200_launch(c, flags, nullptr);
    if (rc != FEPB200_OK)
    {
        return rc;
    }
    if (c->lap_on)
    {
        cudaEventRecord(c->lap_ev[2], c->stream);
    }
    lap_us(c, 2);
    rc = fepb200_download(c, flags, f, fshift, Vc, Vv, dvdl, foreign_energy, foreign_dvdl);
    if (c->lap_on && rc == FEPB200_OK)
    {
        for (int i = 0; i < 3; i++)
        {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, c->lap_ev[i], c->lap_ev[i + 1]) == cudaSuccess)
            {
                c->lap_dev[i] += 1e3 * ms;
            }
        }
    }
    return rc;
}

int fepb200_set_profiling(fepb200_ctx* c, int on)
{
    if (!c)
    {
        return FEPB200_ERR_INVALID_ARGUMENT;
    }
    cudaSetDevice(c->device);
    if (on && !c->ev_prof[0])
    {
        for (int i = 0; i < 4; i++)
        {
            CU_CHECK(c, cudaEventCreate(&c->ev_prof[i]));
        }
    }
    c->profiling = on != 0;
    return FEPB200_OK;
}

int fepb200_kernel_ms(fepb200_ctx* c, float* ms)
{
    if (!c || !ms || !c->profiled)
    {
        return fail(c, FEPB200_ERR_STATE, "the last launch was not profiled (fepb200_set_profiling)");
    }
    cudaSetDevice(c->device);
    CU_CHECK(c, cudaEventSynchronize(c->ev_prof[3]));
    for (int i = 0; i < 3; i++)
    {
        CU_CHECK(c, cudaEventElapsedTime(&ms[i], c->ev_prof[i], c->ev_prof[i + 1]));
    }
    return FEPB200_OK;
}

long long fepb200_launch_count(const fepb200_ctx* c)
{
    return c ? c->launches : 0;
}

int fepb200_last_launch_ms(fepb200_ctx* c, float* ms)
{
    if (!c || !ms || !c->timed)
    {
        return fail(c, FEPB200_ERR_STATE, "no launch has been timed");
    }
    cudaSetDevice(c->device);
    close_chain(c);
    CU_CHECK(c, cudaEventSynchronize(c->ev_stop));
    CU_CHECK(c, cudaEventElapsedTime(ms, c->ev_start, c->ev_stop));
    return FEPB200_OK;
}


/* =========================================================================================== */
/* Perturbed 1-4 pair interactions.
 *
 * Replaces the perturbed branch of do_pairs(F_LJ14, ...) (listed_forces/pairs.cpp:516-835 with
 * free_energy_evaluate_single, :170-515) and the fork's pairs_fep_gpu
 * (listed_forces/listed_forces_gpu_internal.cu:1365-1500).  The soft-core mathematics is the one of
 * the non-bonded perturbed pairs without cut-offs, shifts and long-range corrections, so the
 * pairs are run through the same kernels: every 1-4 pair becomes an i-entry with one j atom of a
 * private pair list whose "atoms" are per-pair copies (i copy: charges of ai and the 1-4 type as
 * its A/B type; j copy: charges of aj), evaluated with plain Coulomb (epsfac * fudgeQQ, k_rf = c_rf
 * = 0), plain LJ and effectively infinite cut-offs.  The minimum-image shift of pbc_dx_aiuc
 * (pbcutil/pbc.cpp:825-851, rectangular boxes) is applied to the i copy's coordinates on the host
 * each step, and the shift forces follow from the per-pair forces (pairs.cpp:822-829). */
/* =========================================================================================== */
struct fepb200_pairs14
{
    fepb200_ctx*       ctx = nullptr;
    std::string        error;
    fepb200_params     ic{};
    float              fudge       = 1.0f;
    bool               have_params = false, have_pairs = false;
    int                natoms = 0, npairs = 0, ngrp = 1;
    std::vector<int>   ai, aj, shift_idx;
    std::vector<float> xp, fp, zero_shift;
    std::vector<double> vc, vv;
    float              last_lambda[FEPB200_NUM_LAMBDA_COMPONENTS];
    bool               have_lambda = false;
};

static int fail14(fepb200_pairs14* h, int code, const char* msg)
{
    if (h)
    {
        h->error = msg;
    }
    else
    {
        g_create_error = msg;
    }
    return code;
}

static int inner14(fepb200_pairs14* h, int rc)
{
    if (rc != FEPB200_OK)
    {
        h->error = fepb200_last_error(h->ctx);
    }
    return rc;
}

int fepb200_pairs14_create(fepb200_pairs14** out, int device_ordinal)
{
    if (!out)
    {
        return fail14(nullptr, FEPB200_ERR_INVALID_ARGUMENT, "handle pointer is NULL");
    }
    *out                = nullptr;
    fepb200_pairs14* h  = new fepb200_pairs14();
    const int        rc = fepb200_create(&h->ctx, device_ordinal);
    if (rc != FEPB200_OK)
    {
        delete h;
        return rc; /* message already in the create-error slot */
    }
    h->zero_shift.assign(3 * FEP_NUM_SHIFT, 0.0f);
    *out = h;
    return FEPB200_OK;
}

int fepb200_pairs14_destroy(fepb200_pairs14* h)
{
    if (h)
    {
        fepb200_destroy(h->ctx);
        delete h;
    }
    return FEPB200_OK;
}

const char* fepb200_pairs14_last_error(const fepb200_pairs14* h)
{
    return h ? h->error.c_str() : g_create_error.c_str();
}

int fepb200_pairs14_set_params(fepb200_pairs14* h, const fepb200_params* ic, float fudgeQQ)
{
    if (!h || !ic)
    {
        return fail14(h, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_pairs14_set_params: NULL argument");
    }
    fepb200_params p            = *ic;
    p.eeltype                   = FEPB200_EEL_CUT; /* plain Coulomb: reaction-field form with k_rf = c_rf = 0 */
    p.vdwtype                   = FEPB200_VDW_CUT;
    p.vdw_modifier              = FEPB200_MOD_NONE;
    p.epsfac                    = ic->epsfac * fudgeQQ; /* pairs.cpp:624 */
    p.rcoulomb                  = 1.0e15f;              /* no cut-off for 1-4 pairs */
    p.rvdw                      = 1.0e15f;
    p.rvdw_switch               = 0.0f;
    p.reactionFieldCoefficient  = 0.0f;
    p.reactionFieldShift        = 0.0f;
    p.sh_ewald                  = 0.0f;
    p.sh_lj_ewald               = 0.0f;
    p.dispersion_shift_cpot     = 0.0f;
    p.repulsion_shift_cpot      = 0.0f;
    const int rc                = inner14(h, fepb200_set_params(h->ctx, &p));
    if (rc != FEPB200_OK)
    {
        return rc;
    }
    /* the Gapsys linearisation point uses the un-fudged epsfac and the real Coulomb cut-off */
    h->ctx->ka.gapsys_facel = ic->epsfac;
    h->ctx->ka.gapsys_rcoul = ic->rcoulomb;
    h->ic                   = *ic;
    h->fudge                = fudgeQQ;
    h->have_params          = true;
    return FEPB200_OK;
}

int fepb200_pairs14_set_pairs(fepb200_pairs14* h, int natoms, const float* chargeA, const float* chargeB, int npairs,
                              const int* iatoms, int ntypes, const float* c6A, const float* c12A, const float* c6B,
                              const float* c12B, const int* gid, int nenergrp_pairs)
{
    if (!h || natoms < 0 || npairs < 0 || ntypes < 1 || nenergrp_pairs < 1
        || (npairs > 0 && (!chargeA || !chargeB || !iatoms || !c6A || !c12A || !c6B || !c12B)))
    {
        return fail14(h, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_pairs14_set_pairs: bad arguments");
    }
    if (!h->have_params)
    {
        return fail14(h, FEPB200_ERR_STATE, "fepb200_pairs14_set_params() has not been called");
    }
    if (ntypes > 2048)
    {
        return fail14(h, FEPB200_ERR_UNSUPPORTED, "more than 2048 perturbed 1-4 interaction types");
    }
    /* 1-4 type t -> inner atom types 2t (state A) and 2t+1 (state B) of the i copy; j copies have
     * type 0, so the pair parameters sit in column 0 of the inner table */
    const int          nt = 2 * ntypes;
    std::vector<float> nbfp(2 * (size_t)nt * nt, 0.0f);
    for (int t = 0; t < ntypes; t++)
    {
        nbfp[2 * ((size_t)nt * (2 * t) + 0)]         = 6.0f * c6A[t]; /* pairs.cpp:655-656 */
        nbfp[2 * ((size_t)nt * (2 * t) + 0) + 1]     = 12.0f * c12A[t];
        nbfp[2 * ((size_t)nt * (2 * t + 1) + 0)]     = 6.0f * c6B[t]; /* pairs.cpp:686-687 */
        nbfp[2 * ((size_t)nt * (2 * t + 1) + 0) + 1] = 12.0f * c12B[t];
    }
    int rc = inner14(h, fepb200_set_nbfp(h->ctx, nt, nbfp.data(), nullptr));
    if (rc != FEPB200_OK)
    {
        return rc;
    }
    std::vector<float> qa(2 * (size_t)npairs), qb(2 * (size_t)npairs);
    std::vector<int>   ta(2 * (size_t)npairs), tb(2 * (size_t)npairs), iinr(npairs), shift(npairs, FEP_CENTRAL_SHIFT),
            jindex(npairs + 1), jjnr(npairs), gids(npairs, 0);
    h->ai.resize(npairs);
    h->aj.resize(npairs);
    for (int p = 0; p < npairs; p++)
    {
        const int t = iatoms[3 * p], i = iatoms[3 * p + 1], j = iatoms[3 * p + 2];
        if (t < 0 || t >= ntypes || i < 0 || i >= natoms || j < 0 || j >= natoms || (gid && (gid[p] < 0 || gid[p] >= nenergrp_pairs)))
        {
            return fail14(h, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_pairs14_set_pairs: malformed pair");
        }
        h->ai[p]      = i;
        h->aj[p]      = j;
        qa[2 * p]     = chargeA[i];
        qb[2 * p]     = chargeB[i];
        ta[2 * p]     = 2 * t;
        tb[2 * p]     = 2 * t + 1;
        qa[2 * p + 1] = chargeA[j];
        qb[2 * p + 1] = chargeB[j];
        ta[2 * p + 1] = 0;
        tb[2 * p + 1] = 0;
        iinr[p]       = 2 * p;
        jjnr[p]       = 2 * p + 1;
        jindex[p]     = p;
        gids[p]       = gid ? gid[p] : 0;
    }
    jindex[npairs] = npairs;
    if ((rc = inner14(h, fepb200_set_atoms(h->ctx, 2 * npairs, qa.data(), qb.data(), ta.data(), tb.data()))) != FEPB200_OK
        || (rc = inner14(h, fepb200_set_list(h->ctx, npairs, iinr.data(), gids.data(), shift.data(), jindex.data(),
                                             jjnr.data(), nullptr, nenergrp_pairs, 0, 1)))
                   != FEPB200_OK)
    {
        return rc;
    }
    h->natoms = natoms;
    h->npairs = npairs;
    h->ngrp   = nenergrp_pairs;
    h->xp.assign(6 * (size_t)npairs, 0.0f);
    h->fp.assign(6 * (size_t)npairs, 0.0f);
    h->shift_idx.assign(npairs, FEP_CENTRAL_SHIFT);
    h->vc.assign(nenergrp_pairs, 0.0);
    h->vv.assign(nenergrp_pairs, 0.0);
    h->have_pairs = true;
    return FEPB200_OK;
}

/* per-pair copies of the coordinates; the minimum-image shift goes onto the i copy (pbc_dx_aiuc, pbcutil/pbc.cpp:825-851) */
static void stage_pairs14(fepb200_pairs14* h, const float* x, const float* box_diag, int pbc_type)
{
    for (int p = 0; p < h->npairs; p++)
    {
        const float* xi = x + 3 * (size_t)h->ai[p];
        const float* xj = x + 3 * (size_t)h->aj[p];
        int          is[3] = { 0, 0, 0 };
        for (int d = 0; d < 3; d++)
        {
            float sh = 0.0f;
            if (pbc_type == 1 || (pbc_type == 2 && d < 2))
            {
                const float dx = xi[d] - xj[d], hbox = 0.5f * box_diag[d];
                if (dx > hbox)
                {
                    sh = -box_diag[d];
                    is[d]--;
                }
                else if (dx <= -hbox)
                {
                    sh = box_diag[d];
                    is[d]++;
                }
            }
            h->xp[6 * (size_t)p + d]     = xi[d] + sh;
            h->xp[6 * (size_t)p + 3 + d] = xj[d];
        }
        h->shift_idx[p] = 5 * (3 * (is[2] + 1) + (is[1] + 1)) + (is[0] + 2);
    }
}

/* All foreign lambda points of the perturbed 1-4 pairs in ONE evaluation: the reference calls the pair code once per
 * point (calc_listed_lambda inside the loop of ListedForces::calculate, listed_forces/listed_forces.cpp:760-800); here
 * the points go through the foreign-lambda machinery of the non-bonded kernels (one load of every pair, the
 * lambda-independent part evaluated once).  energy[i] = Coulomb-14 + LJ-14 energy at point i summed over the
 * energy-group pairs (what sum_epot makes of them), dvdl[2 i + {0, 1}] = dV/dlambda coul / vdw at point i (energy-only
 * semantics, like the reference's foreign evaluations).  Stored, not accumulated. */
int fepb200_pairs14_compute_foreign(fepb200_pairs14* h, const float* x, const float* box_diag, int pbc_type, int n_points,
                                    const float* lambda_coul, const float* lambda_vdw, double* energy, double* dvdl)
{
    if (!h || !x || n_points < 1 || !lambda_coul || !lambda_vdw || !energy || !dvdl || (pbc_type != 0 && !box_diag)
        || pbc_type < 0 || pbc_type > 2)
    {
        return fail14(h, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_pairs14_compute_foreign: bad arguments");
    }
    if (!h->have_pairs)
    {
        return fail14(h, FEPB200_ERR_STATE, "fepb200_pairs14_set_pairs() has not been called");
    }
    for (int i = 0; i < n_points; i++)
    {
        energy[i] = dvdl[2 * i] = dvdl[2 * i + 1] = 0.0;
    }
    if (h->npairs == 0)
    {
        return FEPB200_OK;
    }
    /* the points as the foreign list of the private context (its "current lambda" = the first point; the current-lambda
     * pass is not run) */
    float lam[FEPB200_NUM_LAMBDA_COMPONENTS] = {};
    lam[FEPB200_LAMBDA_COUL]                 = lambda_coul[0];
    lam[FEPB200_LAMBDA_VDW]                  = lambda_vdw[0];
    int rc = inner14(h, fepb200_set_lambdas(h->ctx, lam, n_points, lambda_coul, lambda_vdw));
    h->have_lambda = false; /* fepb200_pairs14_compute() must set its own lambda again */
    if (rc != FEPB200_OK)
    {
        return rc;
    }
    stage_pairs14(h, x, box_diag, pbc_type);
    std::vector<double> fe((size_t)n_points + 1), fd(2 * ((size_t)n_points + 1));
    double              dv[2] = { 0.0, 0.0 };
    rc = inner14(h, fepb200_compute(h->ctx, h->xp.data(), h->zero_shift.data(), FEPB200_DO_FOREIGNLAMBDA | FEPB200_CLEAR_OUTPUTS,
                                    nullptr, nullptr, nullptr, nullptr, dv, fe.data(), fd.data()));
    if (rc != FEPB200_OK)
    {
        return rc;
    }
    /* point 0 of the context repeats its current lambda (freeenergydispatch.cpp:247-253); the caller's points follow */
    for (int i = 0; i < n_points; i++)
    {
        energy[i]       = fe[i + 1];
        dvdl[2 * i]     = fd[2 * (i + 1)];
        dvdl[2 * i + 1] = fd[2 * (i + 1) + 1];
    }
    return FEPB200_OK;
}

int fepb200_pairs14_compute(fepb200_pairs14* h, const float* x, const float* box_diag, int pbc_type, const float* lambda,
                            int flags, float* f, float* fshift, double* Vc14, double* Vv14, double* dvdl)
{
    if (!h || !x || !lambda || !dvdl || (pbc_type != 0 && !box_diag) || pbc_type < 0 || pbc_type > 2)
    {
        return fail14(h, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_pairs14_compute: bad arguments");
    }
    if (!h->have_pairs)
    {
        return fail14(h, FEPB200_ERR_STATE, "fepb200_pairs14_set_pairs() has not been called");
    }
    const bool do_f = (flags & FEPB200_DO_FORCE) != 0, do_vir = do_f && (flags & FEPB200_DO_SHIFTFORCE) != 0,
               do_e = (flags & FEPB200_DO_POTENTIAL) != 0;
    if ((do_f && !f) || (do_vir && !fshift) || (do_e && (!Vc14 || !Vv14)))
    {
        return fail14(h, FEPB200_ERR_INVALID_ARGUMENT, "fepb200_pairs14_compute: an output requested by flags is NULL");
    }
    int rc;
    if (!h->have_lambda || std::memcmp(h->last_lambda, lambda, sizeof(h->last_lambda)) != 0)
    {
        if ((rc = inner14(h, fepb200_set_lambdas(h->ctx, lambda, 0, nullptr, nullptr))) != FEPB200_OK)
        {
            return rc;
        }
        std::memcpy(h->last_lambda, lambda, sizeof(h->last_lambda));
        h->have_lambda = true;
    }
    if (h->npairs == 0)
    {
        return FEPB200_OK;
    }
    stage_pairs14(h, x, box_diag, pbc_type);
    const int inner_flags = (flags & (FEPB200_DO_FORCE | FEPB200_DO_POTENTIAL)) | FEPB200_CLEAR_OUTPUTS;
    double    dv[2]       = { 0.0, 0.0 };
    rc = inner14(h, fepb200_compute(h->ctx, h->xp.data(), h->zero_shift.data(), inner_flags, h->fp.data(), nullptr,
                                    h->vc.data(), h->vv.data(), dv, nullptr, nullptr));
    if (rc != FEPB200_OK)
    {
        return rc;
    }
    if (do_f)
    {
        for (int p = 0; p < h->npairs; p++)
        {
            const float* fi = h->fp.data() + 6 * (size_t)p;
            float*       fa = f + 3 * (size_t)h->ai[p];
            float*       fb = f + 3 * (size_t)h->aj[p];
            for (int d = 0; d < 3; d++)
            {
                fa[d] += fi[d];
                fb[d] += fi[3 + d];
            }
            if (do_vir && h->shift_idx[p] != FEP_CENTRAL_SHIFT)
            {
                for (int d = 0; d < 3; d++)
                {
                    fshift[3 * h->shift_idx[p] + d] += fi[d];
                    fshift[3 * FEP_CENTRAL_SHIFT + d] -= fi[d];
                }
            }
        }
    }
    if (do_e)
    {
        for (int g = 0; g < h->ngrp; g++)
        {
            Vc14[g] += h->vc[g];
            Vv14[g] += h->vv[g];
        }
    }
    dvdl[0] += dv[0];
    dvdl[1] += dv[1];
    return FEPB200_OK;
}

} /* extern "C" */
