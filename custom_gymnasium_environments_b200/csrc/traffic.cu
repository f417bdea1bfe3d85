// Batched TrafficManagementEnv for sm_100a: light control + light timers + vehicle spawn/route + queue processing +
// reward + termination + auto-reset + the (14*NI+4)-feature observation in ONE kernel.
//
// Reference behaviour (paths relative to the reference root, directory traffic_management_env/):
//   config.py:6-35                      constants
//   utils.py:71-118   TrafficLight      update / _advance_phase / can_pass / set_phase
//   utils.py:121-171  Intersection      queues, process_vehicles
//   utils.py:174-248  routes            generate_vehicle_route, neighbours (N,S,W,E), direction
//   environment.py:141-166 reset, :168-203 step, :205-220 _apply_actions, :222-249 _spawn_vehicles,
//   :251-285 update/process/remove, :287-311 _calculate_reward, :313-363 _get_observation
//
// Vehicles never move in the reference: a vehicle waits in its start intersection's
// queue until the light lets the whole queue go, then either leaves `self.vehicles` (its random route ended where
// it started) or stays there forever.  A queue is therefore exactly (length, sum of waiting times, loop-back
// count) and the env needs one more integer, len(self.vehicles).  All dynamics are integer; reward and
// observation are float64 expressions of those integers in the reference's operation order (np.var = NumPy's
// pairwise sums), cast to float32 -- bit-exact against the reference.
//
// One thread per env over [field][env] arrays (coalesced); the T x obs_dim float observation tile is composed in
// shared memory and drained with one bulk asynchronous copy (cp.async.bulk, UBLKCP).  HBM-bound on paper
// (~1.2 KB per env-step) but at 65,536 envs one step is only ~80 MB, so launch latency matters as much.
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "beng_common.cuh"
#include "beng_rng.cuh"

namespace beng {
namespace {

constexpr int MAXNI = BENG_TRAFFIC_MAX_INTERSECTIONS;
constexpr uint32_t TFLAG_NEEDS_RESET = 1u;
enum { NS_GREEN = 0, NS_YELLOW = 1, EW_GREEN = 2, EW_YELLOW = 3 };
enum { NORTH = 0, EAST = 1, SOUTH = 2, WEST = 3 };

struct TArgs {
    beng_traffic_params p;
    beng_traffic_state st;
    beng_traffic_io io;
    const long long *actions;
    const uint8_t *mask;
    long long n;
    int ni, first_call;
};

// NumPy's pairwise summation (np.var in _calculate_reward), n <= 128
template <int CAP>
__device__ __forceinline__ double np_sum(const double (&a)[CAP], int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += a[i];
        return r;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

// generate_vehicle_route (utils.py:174-193) + get_direction_between_intersections (:230-248) reduced to what the
// dynamics use: the first hop's direction and whether the walk ends on its start.
__device__ __forceinline__ void spawn_route(const TArgs &a, EnvStream &rng, int &start, int &dir, int &loopback) {
    const int rows = a.p.grid_rows, cols = a.p.grid_cols;
    start = rng.randint(0, a.ni - 1);
    const int route_length = rng.randint(2, a.ni < 5 ? a.ni : 5);
    int row = start / cols, col = start % cols;
    dir = NORTH;
    for (int h = 1; h < route_length; ++h) {
        // neighbours in the reference's order N, S, W, E (utils.py:206), bounds of the FULL grid (ids up to rows*cols-1)
        const bool hn = row > 0, hs = row + 1 < rows, hw = col > 0, he = col + 1 < cols;
        const int cnt = hn + hs + hw + he;
        int pick = rng.randint(0, cnt - 1);  // random.choice(neighbors)
        int d;
        if (hn && pick-- == 0) d = NORTH;
        else if (hs && pick-- == 0) d = SOUTH;
        else if (hw && pick-- == 0) d = WEST;
        else d = EAST;
        row += (d == SOUTH) - (d == NORTH);
        col += (d == EAST) - (d == WEST);
        if (h == 1) dir = d;
    }
    loopback = (row * cols + col) == start;
}

template <int NI_T, int T, bool IS_RESET>
__global__ void __launch_bounds__(T) traffic_kernel(const TArgs a) {
    constexpr int CAP = NI_T ? NI_T : MAXNI;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float *tile = reinterpret_cast<float *>(smem_raw);
    const int NI = NI_T ? NI_T : a.ni;
    const int OD = NI * 14 + 4;
    const int tid = threadIdx.x;
    const long long n = a.n;
    const long long first = (long long)blockIdx.x * T;
    const long long env = first + tid;
    const bool active = env < n;
    pdl_launch_dependents();  // the next step's grid may become resident as this one drains ...
    pdl_wait();               // ... and this one touches nothing before the previous step's grid has flushed

    bool ended = false;
    double st_ret = 0.0, st_len = 0.0;
    if (active) {
        float *row = tile + tid * OD;
        uint32_t m0 = a.st.misc[env];
        int timestep = m0 & 0xFFFF;
        uint32_t flags = m0 >> 16;
        int listed = (int)a.st.misc[n + env];
        uint32_t ctr = a.st.misc[2 * n + env];
        double total_reward = a.st.total_reward[env];
        bool do_reset = false, selected = true;
        double rew = 0.0;
        int term = 0;
        if constexpr (IS_RESET) {
            if (a.mask) selected = a.mask[env] != 0;
            do_reset = selected;
            if (selected && a.first_call) ctr = 0;
        } else {
            do_reset = a.p.autoreset_mode == BENG_AUTORESET_NEXT_STEP && (flags & TFLAG_NEEDS_RESET);
        }
        EnvStream rng(a.p.seed, a.p.env_id_base + (uint64_t)env, BENG_STREAM_ENV, ctr);

        // Bring the whole env state into registers with independent, fully coalesced loads issued back to back (one
        // round of memory latency instead of one per intersection); everything below works on these copies.
        uint32_t L[CAP], QM[CAP * 4];
        int P[CAP], Wt[CAP], QW[CAP * 4];
        long long ACT[CAP];
#pragma unroll
        for (int i = 0; i < CAP; ++i) {
            if (i < NI) {
                L[i] = a.st.light[(long long)i * n + env];
                P[i] = a.st.passed[(long long)i * n + env];
                Wt[i] = a.st.waiting[(long long)i * n + env];
                if constexpr (!IS_RESET) ACT[i] = a.actions[env * NI + i];
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    QM[i * 4 + d] = a.st.qmeta[(long long)(i * 4 + d) * n + env];
                    QW[i * 4 + d] = a.st.qwait[(long long)(i * 4 + d) * n + env];
                }
            }
        }

        int sp_i = -1, sp_d = 0, sp_lb = 0;  // vehicle spawned this step: start intersection, direction, loop-back
        if (!IS_RESET && !do_reset) {
            timestep = min(timestep + 1, 65535);  // :170
            // _apply_actions (:205-220) then TrafficLight.update (utils.py:79-97), lights in id order
#pragma unroll
            for (int i = 0; i < CAP; ++i) {
                if (i >= NI) break;
                const uint32_t l = L[i];
                int phase = l & 0xFF, timer = (int)(l >> 8);
                const long long act = ACT[i];
                if (act == 1 && phase != NS_GREEN) { phase = NS_GREEN; timer = 5; }       // set_phase: MIN_PHASE_DURATION
                else if (act == 2 && phase != EW_GREEN) { phase = EW_GREEN; timer = 5; }
                timer -= 1;
                if (timer <= 0) {                                                        // _advance_phase
                    phase = (phase + 1) & 3;
                    timer = (phase & 1) ? 3 : rng.randint(5, 30);                        // YELLOW_DURATION | randint
                }
                const uint32_t nl = (uint32_t)phase | ((uint32_t)timer << 8);
                if (nl != l) a.st.light[(long long)i * n + env] = (uint16_t)nl;
                L[i] = nl;
            }
            // _spawn_vehicles (:222-249)
            if (listed < a.p.max_vehicles) {
                if (rng.random53() < a.p.spawn_rate) {
                    spawn_route(a, rng, sp_i, sp_d, sp_lb);
                    listed += 1;
                }
            }
        }

        // _process_intersections (:271-281, utils.py:141-163), _remove_completed_vehicles (:283-285), and the
        // per-intersection parts of reward (:287-311) and observation (:313-363), one pass over the intersections.
        long long tot_passed = 0, tot_wait = 0, tot_queue = 0;
        double qt[CAP];
#pragma unroll
        for (int i = 0; i < CAP; ++i) {
            if (i < NI) {
                int phase = NS_GREEN, passed = 0, wait = 0;
                if (do_reset) {  // reset (:141-166): fresh TrafficLight (NS_GREEN, timer 0), empty queues, zero counters
                    a.st.light[(long long)i * n + env] = (uint16_t)NS_GREEN;
                    a.st.passed[(long long)i * n + env] = 0;
                    a.st.waiting[(long long)i * n + env] = 0;
                } else {
                    phase = L[i] & 0xFF;
                    passed = P[i];
                    wait = Wt[i];
                }
                const int passed0 = passed, wait0 = wait;
                int qsum = 0;
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    const long long qi = (long long)(i * 4 + d) * n + env;
                    int cnt = 0, lb = 0, qw = 0;
                    uint32_t qm0 = 0;
                    int qw0 = 0;
                    if (do_reset) {
                        a.st.qmeta[qi] = 0;
                        a.st.qwait[qi] = 0;
                    } else {
                        qm0 = QM[i * 4 + d];
                        qw0 = QW[i * 4 + d];
                        cnt = qm0 & 0xFF;
                        lb = qm0 >> 8;
                        qw = qw0;
                        if (!IS_RESET) {
                            if (i == sp_i && d == sp_d) { cnt += 1; lb += sp_lb; }  // add_vehicle_to_queue, waiting_time 0
                            if (cnt) {
                                const bool green = (phase == NS_GREEN && (d == NORTH || d == SOUTH)) ||
                                                   (phase == EW_GREEN && (d == EAST || d == WEST));  // can_pass
                                if (green) {  // the whole queue proceeds; loop-back vehicles leave self.vehicles
                                    passed += cnt;
                                    listed -= lb;
                                    cnt = 0; lb = 0; qw = 0;
                                } else {      // every queued vehicle waits one more step
                                    qw += cnt;
                                    wait += cnt;
                                }
                            }
                            const uint32_t qm = (uint32_t)cnt | ((uint32_t)lb << 8);
                            if (qm != qm0) a.st.qmeta[qi] = (uint16_t)qm;
                            if (qw != qw0) a.st.qwait[qi] = qw;
                        }
                    }
                    qsum += cnt;
                    row[NI * 4 + i * 4 + d] = (float)min(cnt, 20);                          // MAX_QUEUE_LENGTH
                    const double avg = cnt ? (double)qw / (double)cnt : 0.0;
                    row[NI * 8 + i * 4 + d] = (float)(avg < 100.0 ? avg : 100.0);
                }
                if (!do_reset && !IS_RESET) {
                    if (passed != passed0) a.st.passed[(long long)i * n + env] = passed;
                    if (wait != wait0) a.st.waiting[(long long)i * n + env] = wait;
                }
#pragma unroll
                for (int ph = 0; ph < 4; ++ph) row[i * 4 + ph] = (phase == ph) ? 1.0f : 0.0f;
                row[NI * 12 + i * 2] = (float)passed;
                row[NI * 12 + i * 2 + 1] = (float)min(wait, 1000);
                tot_passed += passed;
                tot_wait += wait;
                tot_queue += qsum;
                qt[i] = (double)qsum;
            }
        }

        if (do_reset) {
            timestep = 0;
            total_reward = 0.0;
            listed = 0;
            flags = 0;
        } else if (!IS_RESET) {
            // _calculate_reward (:287-311): CUMULATIVE counters, float64, the reference's order of additions
            rew = 0.0;
            rew += (double)tot_passed * 1.0;
            rew += (double)tot_wait * -0.1;
            rew += (double)tot_queue * -0.05;
            if (NI > 1) {
                const double mean = np_sum(qt, NI) / (double)NI;
                double sq[CAP];
#pragma unroll
                for (int i = 0; i < CAP; ++i) {
                    if (i < NI) { const double dq = qt[i] - mean; sq[i] = dq * dq; }
                }
                const double var = np_sum(sq, NI) / (double)NI;  // np.var: population variance
                rew += 0.5 / (1 + var);
            }
            total_reward += rew;
            term = timestep >= a.p.max_timesteps;  // :196, reported as terminated
        }

        // global metrics (utils.py:251-267, environment.py:352-361)
        const double avg_wait = (double)tot_wait / (double)(tot_passed > 1 ? tot_passed : 1);
        const double avg_queue = (double)tot_queue / (double)NI;
        row[NI * 14 + 0] = (float)listed;
        row[NI * 14 + 1] = (float)(avg_wait < 100.0 ? avg_wait : 100.0);
        row[NI * 14 + 2] = (float)(avg_queue < 50.0 ? avg_queue : 50.0);
        row[NI * 14 + 3] = (float)((double)tot_passed / (double)NI);

        if (!IS_RESET && term && a.p.autoreset_mode != BENG_AUTORESET_DISABLED) {
            ended = true;
            st_ret = total_reward;
            st_len = (double)timestep;
            if (a.io.ep_return) a.io.ep_return[env] = total_reward;
            if (a.io.ep_length) a.io.ep_length[env] = timestep;
            if (a.p.autoreset_mode == BENG_AUTORESET_SAME_STEP) {
                // reset() draws nothing and its observation is a constant: rewrite state + row in place
                for (int i = 0; i < NI; ++i) {
                    a.st.light[(long long)i * n + env] = (uint16_t)NS_GREEN;
                    a.st.passed[(long long)i * n + env] = 0;
                    a.st.waiting[(long long)i * n + env] = 0;
                    for (int d = 0; d < 4; ++d) {
                        a.st.qmeta[(long long)(i * 4 + d) * n + env] = 0;
                        a.st.qwait[(long long)(i * 4 + d) * n + env] = 0;
                    }
                }
                for (int k = 0; k < OD; ++k) row[k] = 0.0f;
                for (int i = 0; i < NI; ++i) row[i * 4] = 1.0f;  // every light NS_GREEN
                timestep = 0;
                total_reward = 0.0;
                listed = 0;
                flags = 0;
            } else {
                flags |= TFLAG_NEEDS_RESET;
            }
        }

        a.st.misc[env] = (uint32_t)timestep | (flags << 16);
        a.st.misc[n + env] = (uint32_t)listed;
        a.st.misc[2 * n + env] = rng.ctr;
        a.st.total_reward[env] = total_reward;
        if constexpr (!IS_RESET) {
            a.io.reward[env] = (float)rew;
            a.io.terminated[env] = (uint8_t)term;
            if (a.io.truncated) a.io.truncated[env] = (uint8_t)(a.p.time_limit_truncation && term);  // raw class: 0, :197
            if (a.io.reward64) a.io.reward64[env] = rew;
        }
    }

    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
        const long long n_here = min((long long)T, n - first);
        const uint32_t bytes = (uint32_t)(n_here * OD * sizeof(float));
        const uint32_t bulk = bytes & ~15u;
        if (bulk) bulk_store_s2g(a.io.obs + first * OD, tile, bulk);
        bulk_commit();
        for (uint32_t i = bulk / 4; i < bytes / 4; ++i) a.io.obs[first * OD + i] = tile[i];  // ragged last tile
    }
    if constexpr (!IS_RESET) {
        if (a.io.stats) {
            const unsigned done_mask = __ballot_sync(0xFFFFFFFFu, ended);
            if (done_mask) {
                double r = st_ret, l = st_len;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    r += __shfl_xor_sync(0xFFFFFFFFu, r, o);
                    l += __shfl_xor_sync(0xFFFFFFFFu, l, o);
                }
                if ((tid & 31) == 0) {
                    atomicAdd(&a.io.stats[0], (double)__popc(done_mask));
                    atomicAdd(&a.io.stats[1], r);
                    atomicAdd(&a.io.stats[2], l);
                }
            }
        }
    }
    if (tid == 0) bulk_wait_read<0>();
}

constexpr int TRAFFIC_T = 64;  // envs (= threads) per CTA; 64 x 520 B = 33 KB observation tile

template <int NI_T, int T, bool IS_RESET>
int launch_t(const TArgs &a, cudaStream_t stream) {
    const size_t smem = (size_t)T * (a.ni * 14 + 4) * sizeof(float);
    const unsigned grid = (unsigned)((a.n + T - 1) / T);
    auto kern = traffic_kernel<NI_T, T, IS_RESET>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = launch_pdl(kern, dim3(grid), dim3(T), smem, stream, a);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return (int)e;
}

template <bool IS_RESET>
int launch(const TArgs &a, cudaStream_t stream) {
    static int tile = -1;  // BENG_TRAFFIC_TILE=32|64 (read once; profiling sweeps)
    if (tile < 0) {
        const char *e = getenv("BENG_TRAFFIC_TILE");
        tile = e ? atoi(e) : TRAFFIC_T;
    }
    if (a.ni == 9) return tile == 32 ? launch_t<9, 32, IS_RESET>(a, stream) : launch_t<9, 64, IS_RESET>(a, stream);
    return launch_t<0, 64, IS_RESET>(a, stream);
}

int check(const beng_traffic_params *p, const beng_traffic_state *st, const beng_traffic_io *io, int64_t n) {
    if (!p || !st || !io || n < 0) return BENG_ERR_BAD_ARG;
    if (!st->light || !st->passed || !st->waiting || !st->qmeta || !st->qwait || !st->misc || !st->total_reward ||
        !io->obs)
        return BENG_ERR_BAD_ARG;
    if ((uintptr_t)io->obs & 15) return BENG_ERR_BAD_ARG;
    if (p->grid_rows < 1 || p->grid_cols < 1 || p->num_intersections < 1) return BENG_ERR_BAD_ARG;
    if (p->autoreset_mode < 0 || p->autoreset_mode > 2) return BENG_ERR_BAD_ARG;
    const long long cells = (long long)p->grid_rows * p->grid_cols;
    const long long ni = p->num_intersections < cells ? p->num_intersections : cells;
    if (ni > MAXNI || p->max_vehicles < 0 || p->max_vehicles > 255) return BENG_ERR_UNSUPPORTED;
    if (p->max_timesteps < 1 || p->max_timesteps > 65535) return BENG_ERR_UNSUPPORTED;
    if (cells < 2) return BENG_ERR_UNSUPPORTED;  // a 1x1 grid has no neighbours to route to
    return 0;
}

TArgs make_args(const beng_traffic_params *p, const beng_traffic_state *st, const beng_traffic_io *io, int64_t n) {
    TArgs a{};
    a.p = *p;
    a.st = *st;
    a.io = *io;
    a.n = n;
    const int cells = p->grid_rows * p->grid_cols;
    a.ni = p->num_intersections < cells ? p->num_intersections : cells;
    return a;
}

}  // namespace
}  // namespace beng

extern "C" {

int beng_traffic_reset(const beng_traffic_params *p, const beng_traffic_state *st, const beng_traffic_io *io,
                       const uint8_t *mask_dev, int64_t n_envs, int32_t first_call, void *stream) {
    if (int rc = beng::check(p, st, io, n_envs)) return rc;
    if (n_envs == 0) return 0;
    beng::TArgs a = beng::make_args(p, st, io, n_envs);
    a.mask = mask_dev;
    a.first_call = first_call;
    return beng::launch<true>(a, (cudaStream_t)stream);
}

int beng_traffic_step(const beng_traffic_params *p, const beng_traffic_state *st, const int64_t *actions_dev,
                      const beng_traffic_io *io, int64_t n_envs, void *stream) {
    if (int rc = beng::check(p, st, io, n_envs)) return rc;
    if (!actions_dev || !io->reward || !io->terminated) return BENG_ERR_BAD_ARG;
    if (n_envs == 0) return 0;
    beng::TArgs a = beng::make_args(p, st, io, n_envs);
    a.actions = (const long long *)actions_dev;
    return beng::launch<false>(a, (cudaStream_t)stream);
}

int beng_traffic_step_host(const beng_traffic_params *p, const beng_traffic_state *st, int64_t *actions_dev,
                           const beng_traffic_io *io, int64_t n_envs, const int64_t *actions_host, float *obs_host,
                           float *reward_host, uint8_t *terminated_host, uint8_t *truncated_host, void *stream) {
    if (!actions_host || !actions_dev) return BENG_ERR_BAD_ARG;
    if (int rc = beng::check(p, st, io, n_envs)) return rc;
    if (truncated_host && !io->truncated) return BENG_ERR_BAD_ARG;
    if (n_envs == 0) return 0;
    const beng::TArgs a = beng::make_args(p, st, io, n_envs);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)n_envs;
    cudaError_t e = cudaMemcpyAsync(actions_dev, actions_host, n * a.ni * sizeof(int64_t), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return (int)e;
    if (int rc = beng_traffic_step(p, st, actions_dev, io, n_envs, stream)) return rc;
#define BENG_D2H(dst, src, bytes) \
    if (dst) { e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s); if (e != cudaSuccess) return (int)e; }
    BENG_D2H(reward_host, io->reward, n * sizeof(float))
    BENG_D2H(terminated_host, io->terminated, n)
    BENG_D2H(truncated_host, io->truncated, n)
    BENG_D2H(obs_host, io->obs, n * (a.ni * 14 + 4) * sizeof(float))
#undef BENG_D2H
    return 0;
}

}  // extern "C"
