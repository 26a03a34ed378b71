// pd_api.cu - the C ABI (include/pd_b200.h): handle management, constant/table upload,
// launch dispatch.  No torch types, no C++ exceptions across the boundary.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/pd_b200.h"
#include "pd_device.cuh"
#include "pd_impl.h"
#include "pd_actor.h"
#include "pd_pso.h"
#include "pd_peak.h"
#include "pd_patch.h"
#include <map>
#include <mutex>

using namespace pd;

static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

static int fail(const std::string &m) {
    g_err = m;
    return 1;
}
#define CK(call)                                                                       \
    do {                                                                               \
        cudaError_t _e = (call);                                                       \
        if (_e != cudaSuccess)                                                         \
            return fail(std::string(#call) + ": " + cudaGetErrorString(_e));           \
    } while (0)

struct PdEnv {
    PdConfig cfg;
    const Impl *impl;
    EnvSoA soa;
    KParams kp;                      // per-handle constants: a __grid_constant__ parameter of every launch
    Scalars<double> &sd = kp.sd;
    Scalars<float> &sf = kp.sf;
    Tables &tb = kp.tb;
    LaunchCtx lc() const { return LaunchCtx{&kp, n_sm, cfg.device}; }
    long long roll_index0 = 0;       // pd_set_rollout_stream: global index of the first local particle
    unsigned int roll_generation = 0;
    std::vector<void *> allocs;
    const double *tape = nullptr;
    int tape_len = 0;
    const double *sigma_uv = nullptr;
    float *wT = nullptr;
    size_t wT_cap = 0;
    int *roll_status = nullptr;
    int *roll_queue = nullptr;
    double *cont_d[2] = {nullptr, nullptr};   // straggler hand-off records A / B (pd_rollout_pso)
    int *cont_i[2] = {nullptr, nullptr};
    int *cont_count = nullptr;       // [2]
    int cont_cap = 0;
    int handoff_steps = 128;         // pd_set_rollout_handoff / pd_set_rollout_stages
    int handoff2_steps = 256;
    int lanes8_below = 0, lanes32_below = 0;   // pd_set_rollout_lanes (0 = default)
    bool handoff_default = true;     // thresholds never set by the caller: per-phase defaults apply
    long long patch_stats[4] = {0, 0, 0, 0};   // aero patches: built / rejected for C_D, for C_L
    std::vector<std::pair<int, uint64_t>> patch_keys_raw;   // (device, signature) of the shared patch sets in use
    double patch_max_err = 0.0;
    // shared-actor collection
    void *w2_img = nullptr;          // bf16 smem image of W2
    const float *w2_src = nullptr;   // which W2 the image was built from
    float *obs_carry = nullptr;      // [n_envs*O] observation the next action is computed from
    bool obs_valid = false;
    unsigned int collect_step = 0;
    int n_sm = 0;                    // queried in pd_create
    bool info_full = false;          // pd_set_info_mode
};

// Every entry point runs on the handle's own device and leaves the caller's current device as it
// found it (a process may hold handles on several GPUs).
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
        if (prev == dev) prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ON_DEVICE(e)                                                        \
    DeviceGuard _guard((e)->cfg.device);                                    \
    if (!_guard.ok) return fail("cannot select the handle's CUDA device")

template <typename T>
static int dev_copy(PdEnv *e, const T *host, size_t n, const T **out) {
    T *d = nullptr;
    CK(cudaMalloc(&d, n * sizeof(T)));
    e->allocs.push_back(d);
    CK(cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice));
    *out = d;
    return 0;
}
template <typename T>
static int dev_alloc(PdEnv *e, size_t n, T **out) {
    T *d = nullptr;
    CK(cudaMalloc(&d, n * sizeof(T)));
    e->allocs.push_back(d);
    CK(cudaMemset(d, 0, n * sizeof(T)));
    *out = d;
    return 0;
}

static int upload_rbf(PdEnv *e, const PdRbfTable &t, RbfDev &d, double *levels_out, int max_points) {
    if (t.n_levels < 1 || t.n_levels > PD_MAX_LEVELS) return fail("rbf table: bad level count");
    if (t.hash_size <= 0 || (t.hash_size & (t.hash_size - 1))) return fail("rbf table: hash size");
    if (t.n_points > max_points) return fail("rbf table: too many data points for the shared staging");
    if (t.n_grids < 1 || t.n_grids > 2) return fail("rbf table: needs 1 or 2 lookup grids");
    d.n_levels = t.n_levels;
    d.n_points = t.n_points;
    d.hash_mask = t.hash_size - 1;
    for (int l = 0; l < 6; ++l) d.off[l] = l <= t.n_levels ? t.level_off[l] : t.level_off[t.n_levels];
    for (int l = 0; l < 5; ++l) levels_out[l] = l < t.n_levels ? t.levels[l] : 1e30;
    if (dev_copy(e, t.mach_sorted, (size_t)t.n_points, &d.mach)) return 1;
    const double *pts = nullptr;
    if (dev_copy(e, t.points, (size_t)t.n_points * 2, &pts)) return 1;
    d.points = reinterpret_cast<const double2 *>(pts);
    const uint8_t *rows = nullptr;
    if (dev_copy(e, t.rows, (size_t)t.n_sets * PD_RBF_ROW_BYTES, &rows)) return 1;
    d.rows = reinterpret_cast<const double *>(rows);
    const unsigned long long *hk = nullptr;
    if (dev_copy(e, (const unsigned long long *)t.hash_keys, (size_t)t.hash_size, &hk)) return 1;
    d.hkeys = hk;
    if (dev_copy(e, t.hash_vals, (size_t)t.hash_size, &d.hvals)) return 1;
    for (int g = 0; g < 2; ++g) {
        const PdRbfGrid &s = t.grids[g < t.n_grids ? g : 0];
        RbfGrid &o = d.grid[g];
        o.m0 = s.m0; o.inv_dm = 1.0 / s.dm; o.a0 = s.a0; o.inv_da = 1.0 / s.da;
        o.nm = s.nm; o.na = s.na;
        if (dev_copy(e, s.cells, (size_t)s.nm * s.na, &o.cells)) return 1;
        const unsigned long long *ih = nullptr;
        static const uint64_t zero64 = 0;
        static const int32_t zero32 = 0;
        if (dev_copy(e, (const unsigned long long *)(s.n_impure ? s.imp_hint : &zero64),
                     (size_t)(s.n_impure ? s.n_impure : 1), &ih)) return 1;
        o.imp_hint = ih;
        if (dev_copy(e, s.n_impure ? s.imp_id : &zero32, (size_t)(s.n_impure ? s.n_impure : 1), &o.imp_id)) return 1;
    }
    return 0;
}

// ---- bicubic patches of the fp32 build (pd_patch.h): one set per (device, table contents), built
// on first use and kept for the life of the process (C_L 4 096 x 512 x (1 x 2) sub-cells: 0.6 GB)
struct PatchKey {
    int device;
    uint64_t sig;
    bool operator<(const PatchKey &o) const { return device != o.device ? device < o.device : sig < o.sig; }
};
struct PatchEntry {
    PatchGridOut out;
    int refs;            // live handles using it; pd_release_aero_patches frees the entries at 0
};
static std::mutex g_patch_mu;
static std::map<PatchKey, PatchEntry> g_patches;

static uint64_t mix_words(uint64_t h, const void *data, size_t bytes) {
    const unsigned char *b = static_cast<const unsigned char *>(data);
    size_t i = 0;
    for (; i + 8 <= bytes; i += 8) {
        uint64_t w;
        memcpy(&w, b + i, 8);
        h = (h ^ w) * 0x9E3779B97F4A7C15ULL;
        h ^= h >> 29;
    }
    for (; i < bytes; ++i) h = (h ^ b[i]) * 0x100000001B3ULL;
    return h;
}

#define PD_PATCH_TOL 1e-8     // absolute, on coefficients of order 1: a sixth of an fp32 ulp at 0.5

static int attach_patches(PdEnv *e, const PdRbfTable &t, RbfDev &d, int g, int sub_x, int sub_y, int which_table) {
    const PdRbfGrid &s = t.grids[g];
    RbfGrid &o = d.grid[g];
    uint64_t sig = 0xCBF29CE484222325ULL;
    sig = mix_words(sig, t.rows, (size_t)t.n_sets * PD_RBF_ROW_BYTES);
    sig = mix_words(sig, t.points, (size_t)t.n_points * 16);
    sig = mix_words(sig, s.cells, (size_t)s.nm * s.na * 4);
    if (s.n_impure) sig = mix_words(sig, s.imp_hint, (size_t)s.n_impure * 8);
    const double hdr[6] = {s.m0, s.dm, s.a0, s.da, (double)(s.nm * 65536 + s.na), (double)(sub_x * 256 + sub_y)};
    sig = mix_words(sig, hdr, sizeof(hdr));
    std::lock_guard<std::mutex> lock(g_patch_mu);
    const PatchKey key{e->cfg.device, sig};
    auto it = g_patches.find(key);
    if (it == g_patches.end()) {
        PatchGridIn in;
        in.rows = d.rows; in.points = d.points; in.cells = o.cells; in.imp_hint = o.imp_hint; in.imp_id = o.imp_id;
        in.cells_host = s.cells; in.imp_hint_host = s.imp_hint;
        in.m0 = s.m0; in.dm = s.dm; in.a0 = s.a0; in.da = s.da; in.nm = s.nm; in.na = s.na;
        in.sub_x = sub_x; in.sub_y = sub_y;
        PatchGridOut out;
        if (build_patch_grid(in, PD_PATCH_TOL, &out, 0)) {
            free_patch_grid(&out);
            return fail(std::string("pd_create: building the aero patches failed: ") + cudaGetErrorString(cudaGetLastError()));
        }
        it = g_patches.emplace(key, PatchEntry{out, 0}).first;
    }
    it->second.refs++;
    e->patch_keys_raw.push_back({key.device, key.sig});
    e->patch_stats[2 * which_table] += it->second.out.n_patches;
    e->patch_stats[2 * which_table + 1] += it->second.out.n_failed;
    e->patch_max_err = std::max(e->patch_max_err, it->second.out.max_err_kept);
    o.pbase = it->second.out.pbase;
    o.patch = it->second.out.patch;
    o.sub_x = sub_x; o.sub_y = sub_y;
    return 0;
}

static int upload_segments(PdEnv *e, const double *x, const double *y, int n, const double **dx,
                           const double **dy, const double **ds, const unsigned char **dlut, double *x0,
                           double *inv_w) {
    if (n < 2 || n > 255) return fail("grid fin table: need 2..255 points");
    if (!(x[n - 1] > x[0])) return fail("grid fin table: abscissae must ascend");
    {   // seg_index_lut: lut[b] = #{x_i < x0 + b w}
        const double range = x[n - 1] - x[0];
        const double w = range / 256.0;
        *x0 = x[0];
        *inv_w = 256.0 / range;
        unsigned char lut[256];
        for (int b = 0; b < 256; ++b) {
            const double start = x[0] + b * w;
            int c = 0;
            while (c < n && x[c] < start) ++c;
            lut[b] = (unsigned char)c;
        }
        if (dev_copy(e, lut, (size_t)256, dlut)) return 1;
    }
    std::vector<double> s(n);
    for (int i = 0; i + 1 < n; ++i) s[i] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
    s[n - 1] = s[n - 2];
    if (dev_copy(e, x, (size_t)n, dx)) return 1;
    if (dev_copy(e, y, (size_t)n, dy)) return 1;
    if (dev_copy(e, s.data(), (size_t)n, ds)) return 1;
    return 0;
}

template <typename R>
static void to_float(const Scalars<double> &a, Scalars<R> &b) {
    const double *src = reinterpret_cast<const double *>(&a);
    R *dst = reinterpret_cast<R *>(&b);
    for (size_t i = 0; i < sizeof(a) / sizeof(double); ++i) dst[i] = (R)src[i];
}

// ICAO-1993 ISA speed of sound on the host (terminal Mach of the supersonic phase); the same
// layer table as the device isa()
static double host_speed_of_sound(double alt) {
    static const double L[8][3] = {{-5.0e3, 320.65, -6.5e-3}, {0.0e3, 288.15, -6.5e-3}, {11.0e3, 216.65, 0.0},
                                   {20.0e3, 216.65, 1.0e-3},  {32.0e3, 228.65, 2.8e-3}, {47.0e3, 270.65, 0.0},
                                   {51.0e3, 270.65, -2.8e-3}, {71.0e3, 214.65, -2.0e-3}};
    if (alt < 0) alt = 0;
    const double RE = 6356766.0;
    const double H = RE * alt / (RE + alt);
    int k = 0;
    for (int j = 0; j < 8; ++j)
        if (H >= L[j][0]) k = j;
    const double T = L[k][1] + L[k][2] * (H - L[k][0]);
    return sqrt(1.4 * 287.05287 * T);
}

static void fill_scalars(const PdConfig &cfg, const PdParams &p, Scalars<double> &s) {
    memset(&s, 0, sizeof(s));
    const double d2r = PD_PI / 180.0, r2d = 180.0 / PD_PI;
    const int n_gim = p.n_engines_gimballed;
    s.T_e = p.thrust_per_engine;
    s.p_e = p.nozzle_exit_pressure;
    s.A_e = p.nozzle_exit_area;
    s.te_over_vex = p.thrust_per_engine / p.v_exhaust;
    if (cfg.phase == PD_PHASE_PURE_THROTTLE) {       // rockets_physics.py:909-934
        s.n_eng = n_gim;
        s.nominal = (0 * 0.4) / n_gim;
        s.dt_phys = 0.025;
    } else if (cfg.phase == PD_PHASE_GIMBALLED) {    // rockets_physics.py:803-836
        s.n_eng = n_gim + 2;
        s.nominal = (3 * 0.4) / n_gim;
        s.dt_phys = 0.1;
    } else {                                         // one Euler step of the env dt, :727-801, 959-997
        s.n_eng = n_gim;
        s.nominal = (cfg.phase == PD_PHASE_SUBSONIC || cfg.phase == PD_PHASE_SUPERSONIC) ? 0.5 : (0 * 0.4) / n_gim;
        s.dt_phys = 0.1;
    }
    // actuator low-pass: the landing burns filter on dt_temp = 0.025 (:805, 911), the flip-over on the env dt
    s.dt_act = cfg.phase == PD_PHASE_FLIP_OVER ? 0.1 : 0.025;
    s.flip_max_gimbal_deg = 10;
    s.one_minus_nominal = 1 - s.nominal;
    s.S_gf = p.grid_fin_area;
    s.d_gf = p.d_base_grid_fin;
    s.R_rocket = p.rocket_radius;
    s.S_ref = p.frontal_area;
    s.m_prop0 = p.propellant_mass_stage1;
    s.c_gust_x = p.c_gust_x;
    s.I_dry = p.inertia[0]; s.h_f = p.inertia[1]; s.h_lower = p.inertia[2]; s.h_ox = p.inertia[3];
    s.m_dry = p.inertia[4]; s.m_f = p.inertia[5]; s.m_ox = p.inertia[6]; s.x_dry = p.inertia[7];
    s.engine_height = p.engine_height;
    s.cop = p.cop;
    const bool ascent = cfg.phase == PD_PHASE_SUBSONIC || cfg.phase == PD_PHASE_SUPERSONIC;
    s.max_gimbal_rad = (ascent ? 7.0 : 5) * d2r;     // math.radians(7.0) :739 / math.radians(5) :831
    s.max_gimbal_deg = (5 * d2r) * r2d;
    s.alive_bonus = 0.01 * (1 - cfg.discount_factor);
    s.speed0 = sqrt(p.initial_state[2] * p.initial_state[2] + p.initial_state[3] * p.initial_state[3]);
    if (p.other) {
        const PdOtherPhases &o = *p.other;
        s.n_eng_ng = o.n_engines_stage1 - n_gim;
        for (int i = 0; i < 13; ++i) s.fi[i] = o.inertia_full[i];
        s.sup_terminal_alt = o.ref_terminal[1];
        s.rcs_force = o.max_rcs_force_per_thruster;
        s.d_rcs_bottom = o.d_base_rcs_bottom;
        s.d_rcs_top = o.d_base_rcs_top;
        if (ascent) {
            s.engine_height = o.engine_height_full;
            s.cop = o.cop_full;
        }
        const int row = cfg.phase == PD_PHASE_SUBSONIC ? 0 : cfg.phase == PD_PHASE_SUPERSONIC ? 1
                        : cfg.phase == PD_PHASE_FLIP_OVER ? 3 : 2;
        for (int i = 0; i < 8; ++i) s.norm8[i] = o.norm_vals[row][i] != 0.0 ? o.norm_vals[row][i] : 1.0;
        // Mach schedules of rtd_rl.py:544-575: max_x, max_vy, max_vx, max_alpha_deg
        static const double SUB[12][5] = {
            {0.0, 50, 10, 10, 0.5}, {0.1, 50, 15, 10, 10}, {0.2, 50, 20, 5, 2}, {0.3, 50, 20, 5, 2},
            {0.4, 50, 20, 5, 2}, {0.5, 50, 20, 5, 2}, {0.6, 50, 20, 5, 1.75}, {0.7, 50, 20, 5, 1.75},
            {0.8, 50, 20, 5, 1.75}, {0.9, 50, 20, 5, 1.75}, {1.0, 50, 20, 5, 1.75}, {1.1, 50, 20, 5, 1.75}};
        static const double SUP[12][5] = {
            {1.0, 100, 50, 9, 8}, {1.1, 100, 60, 20, 8}, {1.5, 100, 60, 20, 8}, {1.75, 100, 60, 30, 8},
            {2.0, 100, 60, 40, 8}, {2.25, 100, 60, 50, 8}, {2.5, 100, 60, 60, 8}, {2.75, 100, 60, 70, 8},
            {3.0, 100, 60, 80, 8}, {3.25, 100, 60, 90, 8}, {3.5, 100, 60, 100, 8}, {3.75, 100, 60, 100, 8}};
        const double (*H)[5] = cfg.phase == PD_PHASE_SUPERSONIC ? SUP : SUB;
        for (int j = 0; j < 12; ++j) {
            s.hyp_m[j] = H[j][0];
            for (int r = 0; r < 4; ++r) s.hyp_v[r][j] = H[j][r + 1];
        }
        for (int r = 0; r < 4; ++r) {
            for (int j = 0; j + 1 < 12; ++j)
                s.hyp_s[r][j] = (s.hyp_v[r][j + 1] - s.hyp_v[r][j]) / (s.hyp_m[j + 1] - s.hyp_m[j]);
            s.hyp_s[r][11] = s.hyp_s[r][10];
        }
        if (cfg.phase == PD_PHASE_SUBSONIC) {
            s.terminal_mach = 1.0;
        } else {        // rtd_rl.py:582-588: Mach of the reference trajectory's last row
            const double yt = o.ref_terminal[1], vxt = o.ref_terminal[2], vyt = o.ref_terminal[3];
            s.terminal_mach = sqrt(vxt * vxt + vyt * vyt) / host_speed_of_sound(yt);
        }
    }
    s.max_defl_rad = 20 * d2r;
    s.norm_y = p.norm_vals[0]; s.norm_vy = p.norm_vals[1];
    s.norm_x = p.norm_vals[5]; s.norm_vx = p.norm_vals[6];
    s.y0 = p.initial_state[1];
    s.mass0 = p.initial_state[8];
    s.k_theta_pso = atanh(0.75) / (25 * d2r);
    s.k_theta_rl = atanh(0.75) / (5 * d2r);
    s.k_thetadot_rl = atanh(0.75) / 0.01;
    s.rl_reward_scale = cfg.rl_reward_scale;
    s.v_opt_a = p.v_opt_a;
    s.v_opt_b = p.v_opt_b;
    static const double L[8][4] = {
        {-5.0e3, 320.65, -6.5e-3, 1.77687e5}, {0.0e3, 288.15, -6.5e-3, 1.01325e5},
        {11.0e3, 216.65, 0.0, 2.26320e4},     {20.0e3, 216.65, 1.0e-3, 5.47487e3},
        {32.0e3, 228.65, 2.8e-3, 8.68014e2},  {47.0e3, 270.65, 0.0, 1.10906e2},
        {51.0e3, 270.65, -2.8e-3, 6.69384e1}, {71.0e3, 214.65, -2.0e-3, 3.95639e0}};
    const double G0 = 9.80665, Rg = 287.05287;
    for (int k = 0; k < 8; ++k) {
        s.isa_Hb[k] = L[k][0]; s.isa_Tb[k] = L[k][1]; s.isa_beta[k] = L[k][2]; s.isa_pb[k] = L[k][3];
        s.isa_boT[k] = L[k][2] / L[k][1];
        s.isa_expo[k] = L[k][2] != 0.0 ? -G0 / (L[k][2] * Rg) : 0.0;
        s.isa_iso[k] = -G0 / (Rg * L[k][1]);
    }
    for (int i = 0; i < 16; ++i) {
        int j = i < p.n_wind ? i : (p.n_wind > 0 ? p.n_wind - 1 : 0);
        s.wind_x[i] = p.n_wind > 0 ? p.wind_alt_km[j] : 0.0;
        s.wind_y[i] = p.n_wind > 0 ? p.wind_speed[j] : 0.0;
    }
    for (int i = 0; i + 1 < p.n_wind && i < 15; ++i)
        s.wind_slope[i] = (p.wind_speed[i + 1] - p.wind_speed[i]) / (p.wind_alt_km[i + 1] - p.wind_alt_km[i]);
    s.log_c[0] = 1.0 / 5.0; s.log_c[1] = -1.0 / 4.0; s.log_c[2] = 1.0 / 3.0; s.log_c[3] = -0.5;
    s.log_c[4] = 0.6931471805599453; s.log_c[5] = 4503599627371519.0; s.log_c[6] = -1.0;
    for (int i = 0; i < 4; ++i) { s.Adu[i] = p.vk_Adu[i]; s.Adv[i] = p.vk_Adv[i]; }
    for (int i = 0; i < 2; ++i) { s.Bdu[i] = p.vk_Bdu[i]; s.Bdv[i] = p.vk_Bdv[i]; }
}

extern "C" {

const char *pd_last_error(void) { return g_err.c_str(); }
int pd_version(void) { return 210; }
uint64_t pd_launch_count(void) { return g_launches.load(); }

int pd_create(const PdConfig *cfg, const PdParams *p, PdEnv **out) {
    if (!cfg || !p || !out) return fail("pd_create: null argument");
    if (cfg->n_envs <= 0) return fail("pd_create: n_envs must be positive");
    if (cfg->phase < 0 || cfg->phase >= PD_N_PHASES) return fail("pd_create: unknown flight phase");
    if (cfg->phase == PD_PHASE_FLIP_OVER && cfg->rtd != PD_RTD_SUPERVISORY)
        return fail("pd_create: flip_over_boostbackburn only works with type='supervisory' upstream (its rl "
                    "and pso truncated_func take one argument, rtd_rl.py:132 / rtd_pso.py:107)");
    if (cfg->phase > PD_PHASE_GIMBALLED && cfg->rtd == PD_RTD_PSO)
        return fail("pd_create: this flight phase only works with type='rl' upstream (its pso "
                    "closures have the wrong arity, rtd_pso.py:38-157)");
    if (cfg->phase > PD_PHASE_GIMBALLED && cfg->phase != PD_PHASE_PCONTROL && !p->other)
        return fail("pd_create: PdParams.other is required for this flight phase");
    if (cfg->rtd < 0 || cfg->rtd > PD_RTD_SUPERVISORY) return fail("pd_create: unknown rtd type");
    if (p->n_wind > 16) return fail("pd_create: wind profile longer than 16 points");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail("pd_create: no CUDA device (this library has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail("pd_create: no such CUDA device");
    PdEnv *e = new PdEnv();
    e->cfg = *cfg;
    ON_DEVICE(e);
    {
        int n_sm = 0;
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, cfg->device) != cudaSuccess || n_sm <= 0) {
            delete e;
            return fail("pd_create: cannot query the SM count");
        }
        e->n_sm = n_sm;
    }
    e->impl = cfg->precision == PD_FP64 ? impl_fp64() : impl_fp32();
    fill_scalars(*cfg, *p, e->sd);
    memset(&e->tb, 0, sizeof(e->tb));
    int rc = upload_rbf(e, p->cd, e->tb.cd, e->sd.cd_levels, 192) || upload_rbf(e, p->cl, e->tb.cl, e->sd.cl_levels, 144);
    rc = rc || upload_segments(e, p->gf_ca_mach, p->gf_ca_val, p->n_gf_ca, &e->tb.ca_x, &e->tb.ca_y, &e->tb.ca_s,
                               &e->tb.ca_lut, &e->tb.ca_x0, &e->tb.ca_inv_w);
    rc = rc || upload_segments(e, p->gf_cn_mach, p->gf_cn_val, p->n_gf_cn, &e->tb.cn_x, &e->tb.cn_y, &e->tb.cn_s,
                               &e->tb.cn_lut, &e->tb.cn_x0, &e->tb.cn_inv_w);
    if (!rc && cfg->precision == PD_FP32 && !cfg->exact_aero) {
        // fp32 production build: bicubic patches of the thin-plate sums (pd_patch.h)
        rc = attach_patches(e, p->cd, e->tb.cd, 0, 4, 8, 0);
        e->tb.cd.grid[1] = e->tb.cd.grid[0];
        rc = rc || attach_patches(e, p->cl, e->tb.cl, 0, 1, 2, 1);
        if (p->cl.n_grids > 1) rc = rc || attach_patches(e, p->cl, e->tb.cl, 1, 2, 1, 1);
        else e->tb.cl.grid[1] = e->tb.cl.grid[0];
    }
    if (rc) { pd_destroy(e); return 1; }
    {   // fast_log table: u_j = double(1/c_j), c_j = 1 + (j + 0.5)/256; second word -log(u_j)
        std::vector<double> tab(2 * PD_LOG_N);
        for (int j = 0; j < PD_LOG_N; ++j) {
            long double cj = 1.0L + ((long double)j + 0.5L) / (long double)PD_LOG_N;
            double u = (double)(1.0L / cj);
            tab[2 * j] = u;
            tab[2 * j + 1] = (double)(-logl((long double)u));
        }
        const double *d = nullptr;
        if (dev_copy(e, tab.data(), tab.size(), &d)) { pd_destroy(e); return 1; }
        e->tb.logtab = reinterpret_cast<const double2 *>(d);
        // SharedTables image: every entry replicated PD_REP times (entry i of copy c at i*PD_REP + c),
        // pulled into shared memory by one TMA bulk copy per block (pd_kernels.cuh: stage_tables)
        std::vector<double> img(pd::PD_SH_IMAGE_BYTES / sizeof(double), 0.0);
        size_t o = 0;
        auto put = [&](const double *src, int n_src, int n_slots) {
            for (int i = 0; i < n_slots; ++i)
                for (int c = 0; c < PD_REP; ++c) {
                    img[o++] = i < n_src ? src[2 * i] : 0.0;
                    img[o++] = i < n_src ? src[2 * i + 1] : 0.0;
                }
        };
        put(tab.data(), PD_LOG_N, PD_LOG_N);
        put(p->cd.points, p->cd.n_points, 256);
        put(p->cl.points, p->cl.n_points, 144);
        const double *di = nullptr;
        if (dev_copy(e, img.data(), img.size(), &di)) { pd_destroy(e); return 1; }
        e->tb.sh_image = di;
    }
    to_float(e->sd, e->sf);
    e->tb.n_ca = p->n_gf_ca;
    e->tb.n_cn = p->n_gf_cn;
    e->tb.n_wind = p->n_wind;
    for (int k = 0; k < 11; ++k) e->tb.init[k] = p->initial_state[k];
    if (cfg->phase >= PD_PHASE_SUBSONIC && cfg->phase <= PD_PHASE_BALLISTIC_ARC)
        for (int k = 0; k < 11; ++k) e->tb.init[k] = p->other->initial_state[cfg->phase - PD_PHASE_SUBSONIC][k];
    if (cfg->phase == PD_PHASE_FLIP_OVER)
        for (int k = 0; k < 11; ++k) e->tb.init[k] = p->other->initial_state[3][k];
    if (p->other && p->other->n_ref >= 2 && (cfg->phase == PD_PHASE_SUBSONIC || cfg->phase == PD_PHASE_SUPERSONIC)) {
        // scipy interp1d sorts by x with a stable sort (reference_trajectory_interpolation.py:14-16)
        const PdOtherPhases &o = *p->other;
        const int n = o.n_ref;
        std::vector<int> ord(n);
        for (int i = 0; i < n; ++i) ord[i] = i;
        std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return o.ref_y[a] < o.ref_y[b]; });
        std::vector<double> y(n), v(n), sl(n);
        for (int i = 0; i < n; ++i) y[i] = o.ref_y[ord[i]];
        if (dev_copy(e, y.data(), (size_t)n, &e->tb.ref_y)) { pd_destroy(e); return 1; }
        const double *src[3] = {o.ref_x, o.ref_vx, o.ref_vy};
        for (int c = 0; c < 3; ++c) {
            for (int i = 0; i < n; ++i) v[i] = src[c][ord[i]];
            for (int i = 0; i + 1 < n; ++i) sl[i] = (v[i + 1] - v[i]) / (y[i + 1] - y[i]);
            sl[n - 1] = sl[n - 2];
            if (dev_copy(e, v.data(), (size_t)n, &e->tb.ref_v[c]) || dev_copy(e, sl.data(), (size_t)n, &e->tb.ref_s[c])) {
                pd_destroy(e);
                return 1;
            }
        }
        e->tb.n_ref = n;
    }
    const size_t B = (size_t)cfg->n_envs;
    EnvSoA &s = e->soa;
    s.n = cfg->n_envs;
    rc = dev_alloc(e, 11 * B, &s.st) || dev_alloc(e, 10 * B, &s.gwin) || dev_alloc(e, B, &s.gwin_n) ||
         dev_alloc(e, 3 * B, &s.aprev) || dev_alloc(e, 6 * B, &s.wst) || dev_alloc(e, B, &s.wctr) ||
         dev_alloc(e, B, &s.episode) ||
         dev_alloc(e, B, &s.trunc_id) || dev_alloc(e, B, &s.ep_steps) || dev_alloc(e, (size_t)1, &s.status) ||
         dev_alloc(e, (size_t)1, &e->roll_status) || dev_alloc(e, (size_t)1, &e->roll_queue);
    if (rc) { pd_destroy(e); return 1; }
    *out = e;
    if (pd_reset(e, nullptr, nullptr)) { pd_destroy(e); *out = nullptr; return 1; }
    CK(cudaDeviceSynchronize());
    return 0;
}

int pd_destroy(PdEnv *e) {
    if (!e) return 0;
    DeviceGuard guard(e->cfg.device);
    {
        std::lock_guard<std::mutex> lock(g_patch_mu);
        for (const auto &k : e->patch_keys_raw) {
            auto it = g_patches.find(PatchKey{k.first, k.second});
            if (it != g_patches.end() && it->second.refs > 0) it->second.refs--;
        }
    }
    for (void *a : e->allocs) cudaFree(a);
    if (e->wT) cudaFree(e->wT);
    for (int k = 0; k < 2; ++k) {
        if (e->cont_d[k]) cudaFree(e->cont_d[k]);
        if (e->cont_i[k]) cudaFree(e->cont_i[k]);
    }
    delete e;
    return 0;
}

// the supervisory closures ride on the rl kernels (StepIO.supervisory overrides their verdict)
static int kernel_rtd(const PdEnv *e) { return e->cfg.rtd == PD_RTD_SUPERVISORY ? PD_RTD_RL : e->cfg.rtd; }

static WindCtx wind_ctx(const PdEnv *e) {
    WindCtx wc;
    wc.tape = e->tape;
    wc.tape_len = e->tape_len;
    wc.seed = e->cfg.seed;
    wc.stochastic = e->cfg.stochastic_wind;
    wc.id_offset = 0;
    return wc;
}

int pd_set_rollout_handoff(PdEnv *e, int steps) {
    if (!e) return fail("pd_set_rollout_handoff: null handle");
    if (steps < 0) return fail("pd_set_rollout_handoff: steps must be >= 0 (0 = off)");
    e->handoff_steps = steps;
    e->handoff2_steps = 2 * steps;
    e->handoff_default = false;
    return 0;
}

int pd_set_rollout_stages(PdEnv *e, int steps, int steps2) {
    if (!e) return fail("pd_set_rollout_stages: null handle");
    if (steps < 0 || steps2 < 0) return fail("pd_set_rollout_stages: steps must be >= 0 (0 = off)");
    e->handoff_steps = steps;
    e->handoff2_steps = steps2 > steps ? steps2 : 2 * steps;
    e->handoff_default = false;
    return 0;
}

int pd_set_rollout_lanes(PdEnv *e, int lanes8_below, int lanes32_below) {
    if (!e) return fail("pd_set_rollout_lanes: null handle");
    if (lanes8_below < 0 || lanes32_below < 0 || (lanes8_below > 0 && lanes32_below > lanes8_below))
        return fail("pd_set_rollout_lanes: need 0 <= lanes32_below <= lanes8_below (0 = default)");
    e->lanes8_below = lanes8_below;
    e->lanes32_below = lanes32_below;
    return 0;
}

int64_t pd_release_aero_patches(void) {
    std::lock_guard<std::mutex> lock(g_patch_mu);
    int64_t freed = 0;
    for (auto it = g_patches.begin(); it != g_patches.end();) {
        if (it->second.refs == 0) {
            DeviceGuard guard(it->first.device);
            freed += (int64_t)it->second.out.n_patches * 64;
            free_patch_grid(&it->second.out);
            it = g_patches.erase(it);
        } else {
            ++it;
        }
    }
    return freed;
}

int pd_aero_patch_stats(PdEnv *e, int64_t *counts, double *max_err) {
    if (!e || !counts) return fail("pd_aero_patch_stats: null argument");
    for (int k = 0; k < 4; ++k) counts[k] = e->patch_stats[k];
    if (max_err) *max_err = e->patch_max_err;
    return 0;
}

int pd_set_info_mode(PdEnv *e, int full) {
    if (!e) return fail("pd_set_info_mode: null handle");
    if (full && (e->cfg.precision != PD_FP64 || e->cfg.enable_wind))
        return fail("pd_set_info_mode: the full info row needs the PD_FP64 build without wind");
    e->info_full = full != 0;
    return 0;
}

int pd_reset(PdEnv *e, const uint8_t *mask, void *stream) {
    if (!e) return fail("pd_reset: null handle");
    ON_DEVICE(e);
    e->impl->reset(e->lc(), e->soa, mask, wind_ctx(e), e->sigma_uv, (cudaStream_t)stream);
    e->obs_valid = false;
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

int pd_step(PdEnv *e, const void *actions, int action_dtype, void *obs, void *reward, uint8_t *done,
            uint8_t *truncated, int32_t *trunc_id, void *next_obs, double *dbg, void *stream) {
    if (!e || !actions) return fail("pd_step: null argument");
    if (action_dtype != PD_ACT_F64 && action_dtype != PD_ACT_F32) return fail("pd_step: action dtype");
    ON_DEVICE(e);
    StepIO io;
    io.actions = actions; io.action_dtype = action_dtype; io.obs = obs; io.reward = reward;
    io.raw_actions = e->cfg.raw_actions || e->cfg.rtd == PD_RTD_SUPERVISORY;
    io.dbg_full = (dbg && e->info_full) ? 1 : 0;
    io.supervisory = e->cfg.rtd == PD_RTD_SUPERVISORY;
    io.next_obs = next_obs; io.done = done; io.truncated = truncated; io.trunc_id = trunc_id;
    io.dbg = dbg;
    e->impl->step(e->lc(), e->cfg.phase, kernel_rtd(e), e->cfg.enable_wind, e->soa, io, wind_ctx(e), e->sigma_uv,
                  e->cfg.auto_reset, (cudaStream_t)stream);
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

int pd_get_state(PdEnv *e, double *state, double *g_window, int32_t *n_window, double *act_prev,
                 void *stream) {
    if (!e) return fail("pd_get_state: null handle");
    ON_DEVICE(e);
    e->impl->get_state(e->soa, state, g_window, n_window, act_prev, (cudaStream_t)stream);
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

int pd_set_state(PdEnv *e, const double *state, const double *g_window, const int32_t *n_window,
                 const double *act_prev, void *stream) {
    if (!e) return fail("pd_set_state: null handle");
    ON_DEVICE(e);
    e->impl->set_state(e->soa, state, g_window, n_window, act_prev, (cudaStream_t)stream);
    e->obs_valid = false;
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

int pd_set_wind_tape(PdEnv *e, const double *tape, int tape_len, const double *sigma_uv) {
    if (!e) return fail("pd_set_wind_tape: null handle");
    e->tape = tape;
    e->tape_len = tape ? tape_len : 0;
    e->sigma_uv = sigma_uv;
    return 0;
}

int pd_pso_update(double *x, double *v, double *best, double *best_fit, const double *fitness,
                  const int32_t *swarm_of, const double *swarm_best, float *weights_out, int n, int P,
                  int64_t index0, double w, double c1, double c2, double lo, double hi, uint64_t seed,
                  int generation, void *stream) {
    if (!x || !v || !best || !best_fit || !fitness || !swarm_of || !swarm_best || n <= 0 || P <= 0)
        return fail("pd_pso_update: bad argument");
    PsoUpdateArgs a;
    a.x = x; a.v = v; a.best = best; a.best_fit = best_fit; a.fitness = fitness; a.swarm_of = swarm_of;
    a.swarm_best = swarm_best; a.weights_out = weights_out; a.n = n; a.P = P; a.index0 = index0;
    a.w = w; a.c1 = c1; a.c2 = c2; a.lo = lo; a.hi = hi; a.seed = seed; a.generation = generation;
    if (pso_update_launch(a, (cudaStream_t)stream)) return fail("pd_pso_update: launch failed");
    g_launches++;
    return 0;
}

int pd_measure_fma_peak(int device, int fp64, double *tflops, double *kernel_ms) {
    if (!tflops) return fail("pd_measure_fma_peak: null argument");
    DeviceGuard guard(device);
    if (!guard.ok) return fail("pd_measure_fma_peak: cannot select the device");
    int n_sm = 0;
    CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    double ms = 0.0;
    if (measure_fma_peak(fp64, n_sm, tflops, &ms)) return fail("pd_measure_fma_peak: kernel failed");
    if (kernel_ms) *kernel_ms = ms;
    g_launches += 4;
    return 0;
}

int pd_pso_seed_mean(const double *fitness, int n, int n_seeds, double *out, void *stream) {
    if (!fitness || !out || n <= 0 || n_seeds <= 0) return fail("pd_pso_seed_mean: bad argument");
    if (pso_seed_mean_launch(fitness, n, n_seeds, out, (cudaStream_t)stream)) return fail("pd_pso_seed_mean: launch failed");
    g_launches++;
    return 0;
}

int pd_pso_select(const double *allfit, const int32_t *swarm_of_all, int N, int S, double *swarm_best_fit,
                  int32_t *sel_idx, int32_t *improved, double *stats, void *stream) {
    if (!allfit || !swarm_of_all || !swarm_best_fit || !sel_idx || !improved || !stats || N <= 0 || S <= 0)
        return fail("pd_pso_select: bad argument");
    PsoSelectArgs a;
    a.allfit = allfit; a.swarm_of_all = swarm_of_all; a.N = N; a.S = S; a.swarm_best_fit = swarm_best_fit;
    a.sel_idx = sel_idx; a.improved = improved; a.stats = stats;
    if (pso_select_launch(a, (cudaStream_t)stream)) return fail("pd_pso_select: launch failed");
    g_launches++;
    return 0;
}

int pd_pso_gather(const double *x, int64_t lo, int n_local, int P, const int32_t *sel_idx,
                  const int32_t *improved, int S, double *cand, void *stream) {
    if (!x || !sel_idx || !improved || !cand || n_local < 0 || P <= 0 || S <= 0)
        return fail("pd_pso_gather: bad argument");
    if (pso_gather_launch(x, lo, n_local, P, sel_idx, improved, S, cand, (cudaStream_t)stream))
        return fail("pd_pso_gather: launch failed");
    g_launches++;
    return 0;
}

int pd_pso_apply(const double *cand, const int32_t *improved, int S, int P, double *swarm_best,
                 const double *swarm_best_fit, double *gbest_pos, double *gbest_fit, double *hist_row,
                 void *stream) {
    if (!cand || !improved || !swarm_best || !swarm_best_fit || !gbest_pos || !gbest_fit || S <= 0 || P <= 0)
        return fail("pd_pso_apply: bad argument");
    if (pso_apply_launch(cand, improved, S, P, swarm_best, swarm_best_fit, gbest_pos, gbest_fit, hist_row,
                         (cudaStream_t)stream))
        return fail("pd_pso_apply: launch failed");
    g_launches++;
    return 0;
}

int pd_set_rollout_stream(PdEnv *e, int64_t index0, uint32_t generation) {
    if (!e) return fail("pd_set_rollout_stream: null handle");
    if (index0 < 0) return fail("pd_set_rollout_stream: index0 must be >= 0");
    e->roll_index0 = index0;
    e->roll_generation = generation;
    return 0;
}

int pd_check_status(PdEnv *e, int32_t *status) {
    if (!e || !status) return fail("pd_check_status: null argument");
    ON_DEVICE(e);
    int a = 0, b = 0;
    CK(cudaMemcpy(&a, e->soa.status, sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&b, e->roll_status, sizeof(int), cudaMemcpyDeviceToHost));
    *status = a | b;
    if (*status) return fail("device status: aero-table neighbourhood outside the enumerated set table "
                             "(bit 0) or selection did not converge (bit 1)");
    return 0;
}

static int ensure_wT(PdEnv *e, size_t n) {
    if (e->wT_cap >= n) return 0;
    if (e->wT) cudaFree(e->wT);
    e->wT = nullptr;
    e->wT_cap = 0;
    CK(cudaMalloc(&e->wT, n * sizeof(float)));
    e->wT_cap = n;
    return 0;
}

int pd_rollout_pso(PdEnv *e, const float *weights, int n_particles, int n_params, int n_seeds,
                   int max_steps, double *fitness, int32_t *steps, int32_t *trunc_id,
                   double *terminal_state, double *traj, float *actions_out, double *rewards,
                   void *stream) {
    if (!e || !weights || !fitness) return fail("pd_rollout_pso: null argument");
    const int expect = e->cfg.phase == PD_PHASE_PURE_THROTTLE ? 249 : 372;
    if (n_params != expect) return fail("pd_rollout_pso: n_params does not match the phase's actor");
    if (n_particles <= 0 || n_seeds <= 0 || max_steps <= 0) return fail("pd_rollout_pso: bad sizes");
    if (e->cfg.rtd != PD_RTD_PSO) return fail("pd_rollout_pso: handle was not created with type='pso'");
    if (e->cfg.phase > PD_PHASE_GIMBALLED) return fail("pd_rollout_pso: landing phases only");
    ON_DEVICE(e);
    if (ensure_wT(e, (size_t)n_particles * n_params)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    // landing_burn: the fitness evaluation always runs the fp64 instantiation.  The phase's pitch
    // channel amplifies a perturbation ~50x per 0.4 s env step, so fp32 rounding decides the length
    // of 8 % of even the well-conditioned episodes (tests/golden/pso_many_G.npz: 92 % agreement with
    // the reference in fp32, 99 % in fp64) - a different optimiser landscape, for 20 % of speed.
    const Impl *impl = e->cfg.phase == PD_PHASE_GIMBALLED ? impl_fp64() : e->impl;
    impl->transpose(weights, e->wT, n_particles, n_params, st);
    g_launches++;
    RolloutIO io;
    memset(&io, 0, sizeof(io));
    io.n_episodes = n_particles * n_seeds;
    io.n_seeds = n_seeds;
    io.max_steps = max_steps;
    io.generation = e->roll_generation;
    io.wT = e->wT;
    io.w_stride = (size_t)n_particles;
    io.ret = fitness; io.steps = steps; io.trunc_id = trunc_id; io.terminal = terminal_state;
    io.traj = traj; io.act_out = actions_out; io.rewards = rewards;
    io.queue = e->roll_queue;
    int h1 = e->handoff_steps, h2 = e->handoff2_steps;
    if (e->handoff_default && e->cfg.phase == PD_PHASE_GIMBALLED) {
        // landing_burn episodes last 7-38 steps: one hand-off after 16 steps straight to the final
        // stage while the swarm is small enough for its tail to matter (8 192 particles x 8 seeds:
        // 2.55 -> 2.16 ms), none in the throughput regime (131 072 episodes and more: the extra
        // launches and the record round trip cost 5-9 %, tools/device_swarm_G_probe.py)
        const long long L = (long long)e->n_sm * 448;
        h1 = 2 * (long long)io.n_episodes <= 3 * L ? 16 : 0;
        h2 = max_steps;
    }
    if (h1 > 0) {
        // continuation records of the staged hand-off (two buffers, the stages ping-pong)
        if (e->cont_cap < io.n_episodes) {
            for (int k = 0; k < 2; ++k) {
                if (e->cont_d[k]) { cudaFree(e->cont_d[k]); cudaFree(e->cont_i[k]); e->cont_d[k] = nullptr; e->cont_i[k] = nullptr; }
                CK(cudaMalloc(&e->cont_d[k], (size_t)PD_CONT_D * io.n_episodes * sizeof(double)));
                CK(cudaMalloc(&e->cont_i[k], (size_t)PD_CONT_I * io.n_episodes * sizeof(int)));
            }
            e->cont_cap = io.n_episodes;
        }
        if (!e->cont_count) {
            CK(cudaMalloc(&e->cont_count, 2 * sizeof(int)));
            e->allocs.push_back(e->cont_count);
        }
        io.handoff_steps = h1;
        io.handoff2_steps = h2;
        io.lanes8_below = e->lanes8_below;
        io.lanes32_below = e->lanes32_below;
        io.cont_d = e->cont_d[0]; io.cont_i = e->cont_i[0]; io.cont_count = e->cont_count; io.cont_cap = e->cont_cap;
        io.out_d = e->cont_d[1]; io.out_i = e->cont_i[1]; io.out_count = e->cont_count + 1; io.out_cap = e->cont_cap;
    }
    WindCtx wc = wind_ctx(e);
    // gust-noise stream id = GLOBAL episode index (particle * n_seeds + seed), so that a windy
    // fitness does not depend on how the swarm is sharded over ranks
    wc.id_offset = (unsigned int)(e->roll_index0 * n_seeds);
    if (impl->rollout(e->lc(), PD_POLICY_MLP, e->cfg.phase, e->cfg.rtd, e->cfg.enable_wind, io, wc,
                         e->sigma_uv, e->roll_status, st))
        return fail("pd_rollout_pso: unsupported configuration");
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

int pd_rollout_policy(PdEnv *e, int policy, const void *actions, int action_dtype, int n_episodes,
                      int max_steps, double *ret, int32_t *steps, int32_t *trunc_id,
                      double *terminal_state, double *traj, double *rewards, void *stream) {
    if (!e) return fail("pd_rollout_policy: null handle");
    if (policy != PD_POLICY_TAPE && policy != PD_POLICY_CLASSICAL) return fail("pd_rollout_policy: policy");
    if (policy == PD_POLICY_TAPE && !actions) return fail("pd_rollout_policy: tape policy needs actions");
    if (policy == PD_POLICY_TAPE && e->cfg.rtd == PD_RTD_SUPERVISORY)
        return fail("pd_rollout_policy: type 'supervisory' is served by pd_step only");
    if (n_episodes <= 0 || max_steps <= 0) return fail("pd_rollout_policy: bad sizes");
    ON_DEVICE(e);
    RolloutIO io;
    memset(&io, 0, sizeof(io));
    io.n_episodes = n_episodes;
    io.n_seeds = 1;
    io.max_steps = max_steps;
    io.generation = e->roll_generation;
    io.actions = actions;
    io.action_dtype = action_dtype;
    io.ret = ret; io.steps = steps; io.trunc_id = trunc_id; io.terminal = terminal_state;
    io.traj = traj; io.rewards = rewards;
    io.queue = e->roll_queue;
    if (e->impl->rollout(e->lc(), policy, e->cfg.phase, e->cfg.rtd, e->cfg.enable_wind, io, wind_ctx(e),
                         e->sigma_uv, e->roll_status, (cudaStream_t)stream))
        return fail("pd_rollout_policy: unsupported configuration (classical controller is "
                    "landing_burn_pure_throttle only)");
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

static int actor_args(PdEnv *e, const PdSharedActor *a, ActorArgs &p, int &use_tc, cudaStream_t st) {
    if (!a->w1 || !a->b1 || !a->w2 || !a->b2 || !a->wm || !a->bm || !a->ws || !a->bs)
        return fail("shared actor: null weight pointer");
    if (a->hidden <= 0) return fail("shared actor: hidden must be positive");
    p.w1 = a->w1; p.b1 = a->b1; p.w2 = a->w2; p.b2 = a->b2; p.wm = a->wm; p.bm = a->bm;
    p.ws = a->ws; p.bs = a->bs;
    p.hidden = a->hidden; p.deterministic = a->deterministic; p.max_action = a->max_action;
    p.seed = a->seed; p.step = e->collect_step;
    use_tc = (a->hidden == 256 && !a->fp32_path) ? 1 : 0;
    p.w2_img = nullptr;
    if (use_tc) {
        if (!e->w2_img) {
            CK(cudaMalloc(&e->w2_img, 256 * 256 * 2));
            e->allocs.push_back(e->w2_img);
        }
        // weights change between collection phases (the learner updates them): rebuild the image
        if (actor_prep_w2(a->w2, e->w2_img, st)) return fail("shared actor: W2 image kernel failed");
        g_launches++;
        e->w2_src = a->w2;
        p.w2_img = e->w2_img;
    }
    return 0;
}

int pd_actor_forward(PdEnv *e, const PdSharedActor *actor, const float *obs, int n, float *act,
                     float *mean_out, void *stream) {
    if (!e || !actor || !obs || !act || n <= 0) return fail("pd_actor_forward: bad argument");
    ON_DEVICE(e);
    const int O = pd::phase_odim(e->cfg.phase), A = pd::phase_adim(e->cfg.phase);
    ActorArgs p;
    int use_tc = 0;
    if (actor_args(e, actor, p, use_tc, (cudaStream_t)stream)) return 1;
    if (actor_launch(O, A, p, obs, act, mean_out, n, use_tc, e->n_sm, (cudaStream_t)stream))
        return fail(std::string("pd_actor_forward: launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    g_launches++;
    return 0;
}

int pd_collect_shared_actor(PdEnv *e, const PdSharedActor *actor, int n_steps, float *obs_out,
                            float *act_out, float *rew_out, uint8_t *done_out, uint8_t *trunc_out,
                            float *next_obs_out, void *stream) {
    if (!e || !actor || !act_out || n_steps <= 0) return fail("pd_collect_shared_actor: bad argument");
    if (e->cfg.precision != PD_FP32) return fail("pd_collect_shared_actor: needs the PD_FP32 build (float obs)");
    if (!e->cfg.auto_reset) return fail("pd_collect_shared_actor: handle must be created with auto_reset");
    if (e->cfg.rtd != PD_RTD_RL) return fail("pd_collect_shared_actor: handle must be created with type='rl'");
    ON_DEVICE(e);
    cudaStream_t st = (cudaStream_t)stream;
    const int O = pd::phase_odim(e->cfg.phase), A = pd::phase_adim(e->cfg.phase);
    const size_t B = (size_t)e->cfg.n_envs;
    if (!e->obs_carry) {
        CK(cudaMalloc(&e->obs_carry, B * O * sizeof(float)));
        e->allocs.push_back(e->obs_carry);
    }
    if (!e->obs_valid) {
        e->impl->observe(e->lc(), e->cfg.phase, kernel_rtd(e), e->soa, e->obs_carry, st);
        g_launches++;
        e->obs_valid = true;
    }
    ActorArgs p;
    int use_tc = 0;
    if (actor_args(e, actor, p, use_tc, st)) return 1;
    for (int t = 0; t < n_steps; ++t) {
        float *obs_t = obs_out ? obs_out + (size_t)t * B * O : e->obs_carry;
        if (obs_out && t == 0)
            CK(cudaMemcpyAsync(obs_t, e->obs_carry, B * O * sizeof(float), cudaMemcpyDeviceToDevice, st));
        float *act_t = act_out + (size_t)t * B * A;
        p.step = e->collect_step++;
        if (actor_launch(O, A, p, obs_t, act_t, nullptr, (int)B, use_tc, e->n_sm, st))
            return fail(std::string("pd_collect_shared_actor: actor launch failed: ") +
                        cudaGetErrorString(cudaGetLastError()));
        g_launches++;
        StepIO io;
        io.actions = act_t; io.action_dtype = PD_ACT_F32; io.raw_actions = 0; io.dbg_full = 0; io.supervisory = 0;
        io.obs = next_obs_out ? next_obs_out + (size_t)t * B * O : nullptr;
        io.reward = rew_out ? rew_out + (size_t)t * B : nullptr;
        io.done = done_out ? done_out + (size_t)t * B : nullptr;
        io.truncated = trunc_out ? trunc_out + (size_t)t * B : nullptr;
        io.trunc_id = nullptr;
        io.dbg = nullptr;
        // post-reset observation feeds the next action
        io.next_obs = (obs_out && t + 1 < n_steps) ? (void *)(obs_out + (size_t)(t + 1) * B * O) : (void *)e->obs_carry;
        e->impl->step(e->lc(), e->cfg.phase, kernel_rtd(e), e->cfg.enable_wind, e->soa, io, wind_ctx(e), e->sigma_uv,
                      e->cfg.auto_reset, st);
        g_launches++;
    }
    CK(cudaGetLastError());
    return 0;
}

}  // extern "C"
