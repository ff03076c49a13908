// QB_SCAN_EXACT at scale: a PARALLEL prefix sum that reproduces np.cumsum's strictly sequential fp64
// rounding bit for bit (resamplers.py:308), so that resample indices are bit-identical to the
// reference's at any particle count.
//
// Idea.  While the running sum s stays inside one binade [2^e, 2^(e+1)) its ulp u = 2^(e-52) is fixed and
// k = s / u is an integer in [2^52, 2^53).  Adding a weight w >= 0 in round-to-nearest-even then is the integer
// map  k -> k + q + r  with q = floor(w / u) and r in {0, 1} decided by the discarded fraction — except on an
// exact tie, where r depends on the PARITY of k + q.  So every element is a map of the form
//      k -> k + a   (k even),      k -> k + b   (k odd)
// and such maps are closed under composition: runs of elements can be combined in any grouping — a monoid, i.e.
// scannable in parallel — as long as no element of the run pushes the sum over 2^(e+1).  Binade crossings are rare
// (the sum doubles ~24 times between the first weight and 1) and are handled by replaying the one chunk that
// contains the crossing sequentially.
//
// Kernel.  G co-resident CTAs (1 per SM, 1024 threads), CTA c owns a contiguous segment, thread t a contiguous
// chunk of C elements of it (256-bit loads/stores: one full sector per thread per access).
//   A  predicted binade of the segment (from the approximate tile prefix the fast scan computes) -> per-thread chunk
//      composites -> block scan -> the CTA publishes its aggregate (a, b, e) as PLAIN, or HARD if the prediction
//      is not unique (the segment may contain a crossing);
//   B  decoupled look-back: combine the PLAIN aggregates of the predecessors back to the nearest published exact
//      VALUE, verifying binade and no-crossing at every hop (a failed hop waits for that CTA's VALUE);
//   C  with the exact start value: PLAIN CTAs verify their own prediction, HARD (or mispredicted) ones resolve
//      their crossings iteratively (first crossing thread replays its chunk in plain fp64, the rest is
//      re-composited in the new binade); the CTA publishes its exact end VALUE;
//   D  every thread replays its chunk from its exact start value with ordinary fp64 adds and writes the CDF.
#include "qb_common.cuh"

namespace qb {

constexpr int EX_THREADS = 1024;
constexpr unsigned long long EX_SAT = 1ULL << 60;  // saturation bound for composite increments
constexpr unsigned long long TWO53 = 1ULL << 53;

enum { EX_EMPTY = 0, EX_PLAIN = 1, EX_HARD = 2, EX_VALUE = 3 };

struct ExDesc {  // one per CTA, 64 bytes
    unsigned long long a, b;  // aggregate composite of the segment under binade e
    double value;             // exact inclusive end value of the segment (valid once value_ready)
    int e;
    int status;
    int value_ready;          // 1 once `value` is valid (a PLAIN CTA keeps its aggregate readable until then)
    int pad[7];
};

struct Comp {
    unsigned long long a, b;
};

__device__ __forceinline__ unsigned long long sat(unsigned long long x) { return x > EX_SAT ? EX_SAT : x; }

// element map of weight w in binade e (ulp 2^(e-52)); w >= 0 finite
__device__ __forceinline__ Comp elem_comp(double w, int e) {
    const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(w));
    const int ef = static_cast<int>((bits >> 52) & 0x7ff);
    unsigned long long m = bits & ((1ULL << 52) - 1);
    int ew;  // w = m * 2^(ew - 52)
    if (ef == 0) {
        ew = -1022;  // subnormal or zero: no implicit bit
    } else {
        m |= (1ULL << 52);
        ew = ef - 1023;
    }
    Comp c;
    if (m == 0) {
        c.a = c.b = 0;
        return c;
    }
    const int shift = e - ew;  // ulp(w's grid) is 2^shift times finer than u
    if (shift <= 0) {          // w is a multiple of u
        const int up = -shift;
        const unsigned long long q = (up >= 8) ? EX_SAT : sat(m << up);
        c.a = c.b = q;
        return c;
    }
    if (shift >= 64) {
        c.a = c.b = 0;
        return c;
    }
    const unsigned long long q = m >> shift;
    const unsigned long long rem = m & ((1ULL << shift) - 1);
    const unsigned long long half = 1ULL << (shift - 1);
    if (rem > half) {
        c.a = c.b = q + 1;
    } else if (rem < half) {
        c.a = c.b = q;
    } else {  // exact tie: round to even
        c.a = q + (q & 1ULL);
        c.b = q + ((q & 1ULL) ^ 1ULL);
    }
    return c;
}

// f then g
__device__ __forceinline__ Comp compose(Comp f, Comp g) {
    Comp r;
    r.a = sat(f.a + ((f.a & 1ULL) ? g.b : g.a));
    r.b = sat(f.b + (((f.b + 1ULL) & 1ULL) ? g.b : g.a));
    return r;
}

__device__ __forceinline__ int binade_of(double v) {  // v > 0 finite: floor(log2 v), subnormals map to -1022
    const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(v));
    const int ef = static_cast<int>((bits >> 52) & 0x7ff);
    return ef == 0 ? -1022 : ef - 1023;
}

// k = v / u_e as an integer; valid when v is in binade e (or is a subnormal with e = -1022)
__device__ __forceinline__ unsigned long long to_k(double v, int e) {
    const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(v));
    const int ef = static_cast<int>((bits >> 52) & 0x7ff);
    unsigned long long m = bits & ((1ULL << 52) - 1);
    if (ef != 0) m |= (1ULL << 52);
    (void)e;
    return m;
}
__device__ __forceinline__ double from_k(unsigned long long k, int e) {  // k < 2^53 (or == 2^53)
    return ldexp(static_cast<double>(k), e - 52);
}

struct ExParams {
    const double* w;
    const double* stats;
    const double* tile_prefix;  // approximate exclusive prefix every SCAN tile (2048 weights), ntiles + 1 entries
    const double* carry;        // NULL, or the exact running sum this scan continues from (a slab of a sharded cloud)
    double* cdf;
    ExDesc* desc;
    unsigned int* ticket;
    int* fallback;  // set to 1 if a negative / NaN / inf weight was seen (caller then runs the sequential kernel)
    int64_t n;
    int64_t seg;    // elements per CTA (multiple of 2048)
    int chunk;      // elements per thread (multiple of 4)
    int ncta;
};

// sequential fp64 replay of one chunk from `start`; optionally writes the CDF
template <bool WRITE>
__device__ __forceinline__ double replay_chunk(const ExParams& p, double inv, int64_t first, int cnt, double start) {
    double s = start;
    for (int j = 0; j < cnt; j += 4) {
        double v[4];
        if (j + 4 <= cnt) {
            asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                         : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3])
                         : "l"(p.w + first + j));
        } else {
            for (int q = 0; q < 4; ++q) v[q] = (j + q < cnt) ? p.w[first + j + q] : 0.0;
        }
        double o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (j + q < cnt) s = s + v[q] * inv;
            o[q] = s;
        }
        if (WRITE) {
            if (j + 4 <= cnt) {
                asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p.cdf + first + j), "d"(o[0]), "d"(o[1]),
                             "d"(o[2]), "d"(o[3])
                             : "memory");
            } else {
                for (int q = 0; q < 4 && j + q < cnt; ++q) p.cdf[first + j + q] = o[q];
            }
        }
    }
    return s;
}

// composite of one chunk under binade e; also flags weights the integer model does not cover
__device__ __forceinline__ Comp chunk_comp(const ExParams& p, double inv, int64_t first, int cnt, int e, bool& bad) {
    Comp c = {0ULL, 0ULL};
    for (int j = 0; j < cnt; j += 4) {
        double v[4];
        if (j + 4 <= cnt) {
            asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                         : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3])
                         : "l"(p.w + first + j));
        } else {
            for (int q = 0; q < 4; ++q) v[q] = (j + q < cnt) ? p.w[first + j + q] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (j + q < cnt) {
                const double wv = v[q] * inv;
                if (!(wv >= 0.0) || wv > 1.0e300) bad = true;
                c = compose(c, elem_comp(wv, e));
            }
        }
    }
    return c;
}

// exclusive block scan of composites over threads [t0, EX_THREADS); threads < t0 contribute the identity.
// Returns the exclusive prefix for this thread; total = inclusive composite of the whole block.
__device__ Comp block_scan_comp(Comp mine, Comp* warp_tot, Comp& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    Comp inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Comp prev;
        prev.a = __shfl_up_sync(0xffffffffu, inc.a, o);
        prev.b = __shfl_up_sync(0xffffffffu, inc.b, o);
        if (lane >= o) inc = compose(prev, inc);
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        Comp t = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            Comp prev;
            prev.a = __shfl_up_sync(0xffffffffu, t.a, o);
            prev.b = __shfl_up_sync(0xffffffffu, t.b, o);
            if (lane >= o) t = compose(prev, t);
        }
        warp_tot[lane] = t;  // inclusive over warps
    }
    __syncthreads();
    total = warp_tot[31];
    // exclusive prefix of this thread = (inclusive of previous warps) then (exclusive within warp)
    Comp excl_in_warp;
    excl_in_warp.a = __shfl_up_sync(0xffffffffu, inc.a, 1);
    excl_in_warp.b = __shfl_up_sync(0xffffffffu, inc.b, 1);
    if (lane == 0) excl_in_warp.a = excl_in_warp.b = 0ULL;
    Comp res = excl_in_warp;
    if (wid > 0) res = compose(warp_tot[wid - 1], excl_in_warp);
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(EX_THREADS, 1) exact_scan_kernel(const __grid_constant__ ExParams p) {
    __shared__ Comp warp_tot[32];
    __shared__ unsigned int cta_id;
    __shared__ double sh_start;     // exact value before the first unresolved thread
    __shared__ int sh_first;        // first unresolved thread (HARD resolution loop)
    __shared__ int sh_cross;        // first thread whose chunk crosses the binade in this round
    __shared__ int sh_e, sh_mode;
    __shared__ int sh_bad;

    const int tid = threadIdx.x;
    if (tid == 0) {
        cta_id = atomicAdd(p.ticket, 1u);  // logical CTA ids in start order: a CTA only waits on started ones
        sh_bad = 0;
    }
    __syncthreads();
    const int c = static_cast<int>(cta_id);
    const double inv = p.stats[QB_STAT_INV_NORM];
    const int64_t seg_first = static_cast<int64_t>(c) * p.seg;
    const int64_t seg_cnt64 = (p.n - seg_first < p.seg) ? (p.n - seg_first) : p.seg;
    const int64_t my_first = seg_first + static_cast<int64_t>(tid) * p.chunk;
    int my_cnt = 0;
    if (my_first < seg_first + seg_cnt64) {
        const int64_t rest = seg_first + seg_cnt64 - my_first;
        my_cnt = static_cast<int>(rest < p.chunk ? rest : p.chunk);
    }
    volatile ExDesc* me = p.desc + c;

    // ---- A: predicted binade from the approximate prefix (tiles of 2048; seg is a multiple of 2048) ----
    if (tid == 0) {
        const int64_t t0 = seg_first / 2048;
        const int64_t t1 = (seg_first + seg_cnt64 + 2047) / 2048;
        const double base = p.carry ? *p.carry : 0.0;
        const double lo = (base + p.tile_prefix[t0]) * (1.0 - 1e-11);
        const double hi = (base + p.tile_prefix[t1]) * (1.0 + 1e-11);
        int mode = EX_HARD, e = 0;
        if (c > 0 && lo > 0.0 && hi < 1.0e300 && binade_of(lo) == binade_of(hi)) {
            mode = EX_PLAIN;
            e = binade_of(lo);
        }
        sh_mode = mode;
        sh_e = e;
    }
    __syncthreads();
    int e = sh_e;
    bool bad = false;
    Comp mine = {0ULL, 0ULL};
    Comp excl = {0ULL, 0ULL}, total = {0ULL, 0ULL};
    if (sh_mode == EX_PLAIN) {
        mine = chunk_comp(p, inv, my_first, my_cnt, e, bad);
        excl = block_scan_comp(mine, warp_tot, total);
        if (tid == 0) {
            me->a = total.a;
            me->b = total.b;
            me->e = e;
            __threadfence();
            me->status = EX_PLAIN;
        }
    } else if (tid == 0) {
        __threadfence();
        me->status = EX_HARD;
    }
    if (bad) sh_bad = 1;

    // ---- B: look-back for the exact start value ----
    if (tid == 0) {
        double v = p.carry ? *p.carry : 0.0;
        if (c > 0) {
            int j = c - 1;
            // walk back to the nearest exact VALUE, skipping over PLAIN aggregates
            while (true) {
                volatile ExDesc* dj = p.desc + j;
                const long long t0 = clock64();
                while (dj->status == EX_EMPTY || (dj->status == EX_HARD && !dj->value_ready)) {
                    if (clock64() - t0 > 4000000000LL) {  // ~2 s: never hang the GPU; the sequential kernel takes over
                        *p.fallback = 1;
                        break;
                    }
                }
                if (*(volatile int*)p.fallback) break;
                __threadfence();
                if (dj->value_ready) {
                    __threadfence();
                    v = dj->value;
                    break;
                }
                --j;  // PLAIN aggregate: keep walking (CTA 0 is always HARD, so j never runs below 0)
            }
            // forward over the PLAIN aggregates j+1 .. c-1, verifying every hop
            for (int q = j + 1; q < c; ++q) {
                volatile ExDesc* dq = p.desc + q;
                bool ok = false;
                if (!dq->value_ready && v > 0.0 && v < 1.0e300) {
                    const int eq = dq->e;
                    if (binade_of(v) == eq) {
                        const unsigned long long k = to_k(v, eq);
                        const unsigned long long incr = (k & 1ULL) ? dq->b : dq->a;
                        if (k + incr < TWO53) {
                            v = from_k(k + incr, eq);
                            ok = true;
                        }
                    }
                }
                if (!ok) {  // misprediction or a crossing inside q (or q is already exact): take q's own VALUE
                    const long long t0 = clock64();
                    while (!dq->value_ready) {
                        if (clock64() - t0 > 4000000000LL) {
                            *p.fallback = 1;
                            break;
                        }
                    }
                    __threadfence();
                    v = dq->value;
                }
            }
        }
        sh_start = v;
        sh_first = 0;
    }
    __syncthreads();

    // ---- C: resolve this segment ----
    double my_start = 0.0;
    bool resolved = false;
    {
        const double v = sh_start;
        if (sh_mode == EX_PLAIN && v > 0.0 && binade_of(v) == e) {
            const unsigned long long k = to_k(v, e);
            const unsigned long long tot = (k & 1ULL) ? total.b : total.a;
            if (k + tot < TWO53) {  // prediction holds for the whole segment
                const unsigned long long mk = k + ((k & 1ULL) ? excl.b : excl.a);
                my_start = from_k(mk, e);
                resolved = true;
                if (tid == 0) {
                    me->value = from_k(k + tot, e);
                    __threadfence();
                    me->value_ready = 1;
                    __threadfence();
                    me->status = EX_VALUE;
                }
            }
        }
    }
    const bool all_plain = __syncthreads_and(resolved ? 1 : 0) != 0;
    if (!all_plain) {
        // HARD / mispredicted: iterate over the binades met inside the segment
        resolved = false;
        while (true) {
            const int t0 = sh_first;
            if (t0 >= EX_THREADS) break;
            const double v = sh_start;
            __syncthreads();
            if (v > 0.0 && v < 1.0e300) {
                const int eb = binade_of(v);
                const unsigned long long k = to_k(v, eb);
                bool b2 = false;
                Comp cm = {0ULL, 0ULL};
                if (tid >= t0) cm = chunk_comp(p, inv, my_first, my_cnt, eb, b2);
                if (b2) sh_bad = 1;
                Comp tot2;
                Comp ex2 = block_scan_comp(cm, warp_tot, tot2);
                // my exact start if nobody before me crossed; my end decides whether I cross
                const unsigned long long sk = k + ((k & 1ULL) ? ex2.b : ex2.a);
                const unsigned long long inc_mine = (sk & 1ULL) ? cm.b : cm.a;
                const bool start_ok = sk < TWO53;
                const bool cross = (tid >= t0) && start_ok && (sk + inc_mine >= TWO53);
                if (tid == 0) sh_cross = EX_THREADS;
                __syncthreads();
                if (cross) atomicMin(&sh_cross, tid);
                __syncthreads();
                const int tc = sh_cross;
                if (tid >= t0 && tid < tc && !resolved) {  // everything before the first crossing chunk is exact
                    my_start = from_k(sk, eb);
                    resolved = true;
                }
                if (tc < EX_THREADS) {
                    if (tid == tc) {  // the crossing chunk: plain sequential fp64 from its exact start
                        my_start = from_k(sk, eb);
                        resolved = true;
                        sh_start = replay_chunk<false>(p, inv, my_first, my_cnt, my_start);
                        sh_first = tc + 1;
                    }
                } else if (tid == 0) {
                    sh_start = from_k(k + ((k & 1ULL) ? tot2.b : tot2.a), eb);
                    sh_first = EX_THREADS;
                }
            } else {
                // running sum still zero (leading zero weights) or not finite: replay thread t0's chunk sequentially
                if (tid == t0) {
                    my_start = v;
                    resolved = true;
                    sh_start = replay_chunk<false>(p, inv, my_first, my_cnt, v);
                    sh_first = t0 + 1;
                }
            }
            __syncthreads();
        }
        if (tid == 0) {
            me->value = sh_start;
            __threadfence();
            me->value_ready = 1;
            __threadfence();
            me->status = EX_VALUE;
        }
    }

    // ---- D: replay my chunk from its exact start, writing the CDF ----
    if (my_cnt > 0) replay_chunk<true>(p, inv, my_first, my_cnt, my_start);
    __syncthreads();
    if (tid == 0 && sh_bad) *p.fallback = 1;
}

// the sequential kernel, run only if the parallel one met a weight outside its model (negative, NaN, inf)
__global__ void exact_scan_fallback_kernel(const double* __restrict__ w, const double* __restrict__ stats, int64_t n,
                                           double* __restrict__ cdf, const int* __restrict__ fallback,
                                           const double* __restrict__ carry) {
    if (*fallback == 0 || threadIdx.x != 0 || blockIdx.x != 0) return;
    const double inv = stats[QB_STAT_INV_NORM];
    double run = carry ? *carry : 0.0;
    for (int64_t i = 0; i < n; ++i) {
        run = (i == 0 && !carry) ? w[0] * inv : run + w[i] * inv;
        cdf[i] = run;
    }
}

size_t exact_scan_workspace_bytes(int64_t n) {
    (void)n;
    return static_cast<size_t>(256) * sizeof(ExDesc) + 256;
}

// `tile_prefix`: ntiles + 1 approximate exclusive prefix values (tile = 2048 weights), already on the device.
int launch_exact_scan(const double* d_w, const double* d_stats, int64_t n, double* d_cdf, const double* tile_prefix,
                      const double* d_carry, void* d_ws, cudaStream_t st) {
    const int sms = sm_count();
    int64_t chunk = (n + static_cast<int64_t>(sms) * EX_THREADS - 1) / (static_cast<int64_t>(sms) * EX_THREADS);
    chunk = ((chunk + 3) / 4) * 4;
    if (chunk < 4) chunk = 4;
    const int64_t seg = chunk * EX_THREADS;  // multiple of 4096, hence of the 2048-weight tiles
    const int ncta = static_cast<int>((n + seg - 1) / seg);
    QB_REQUIRE(ncta <= 256, QB_ERR_INVALID_ARGUMENT, "exact scan: too many CTAs (%d)", ncta);
    unsigned char* base = reinterpret_cast<unsigned char*>(d_ws);
    ExParams p;
    p.w = d_w;
    p.stats = d_stats;
    p.tile_prefix = tile_prefix;
    p.carry = d_carry;
    p.cdf = d_cdf;
    p.ticket = reinterpret_cast<unsigned int*>(base);
    p.fallback = reinterpret_cast<int*>(base + 64);
    p.desc = reinterpret_cast<ExDesc*>(base + 256);
    p.n = n;
    p.seg = seg;
    p.chunk = static_cast<int>(chunk);
    p.ncta = ncta;
    QB_CUDA_CHECK(cudaMemsetAsync(base, 0, 256 + sizeof(ExDesc) * ncta, st));
    exact_scan_kernel<<<ncta, EX_THREADS, 0, st>>>(p);
    QB_CUDA_CHECK(cudaGetLastError());
    exact_scan_fallback_kernel<<<1, 32, 0, st>>>(d_w, d_stats, n, d_cdf, p.fallback, d_carry);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

}  // namespace qb
