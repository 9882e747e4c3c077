// replay_host.inl -- host C++ mirror of the reference's steady-state frame loop for B robots in lock step:
//   Estimator::processIMULeg (estimator.cpp:590-653), processImage NON_LINEAR branch (:655-676, :798-846), vector2double / double2vector
//   (:848-1003), optimization (:1054-1456, every numerical step through the C ABI of this library), slideWindow{,Old,New} (:1460-1677),
//   FeatureManager (featureTracker/feature_manager.cpp: addFeatureCheckParallax, setDepth, getDepthVector, removeFailures, removeOutlier,
//   removeBackShiftDepth, removeFront).
// It is the C++ twin of cerberus_b200/estimator.py (same statements, same order; tests/test_replay.py runs both on the same sequences): the
// Python mirror costs ~8 ms of bookkeeping per robot and frame, which caps a batched replay at a few hundred robot-frames per second; this
// one is bound by the device.  Included at the end of cabi.cu (host code only).  The reference's initialisation (stereo PnP + gyroscope-bias
// alignment, estimator.cpp:700-797) is out of scope: a replay is seeded with WINDOW_SIZE + 1 frames at given states.
#include <list>
#include <map>
#include <cmath>

namespace cerbhost {

enum { W = CERB_WINDOW_SIZE, NFRM = CERB_NUM_FRAMES };
static const double kFocal = 460.0, kMinParallax = 10.0 / 460.0, kInitDepth = 5.0;

struct Obs { double point[3], velocity[2], pointRight[3], velocityRight[2], cur_td; bool is_stereo; };
struct Feature { int id, start_frame; std::vector<Obs> obs; int used_num = 0; double estimated_depth = -1.0; int solve_flag = 0; int endFrame() const { return start_frame + (int)obs.size() - 1; } };

struct Interval {                       // an IMULegIntegrationBase: constructor arguments + sample buffers + result
    bool valid = false, dirty = true, has_result = false;
    CerbIMULegSample first; double ba[3], bg[3], rho[4];
    std::vector<CerbIMULegSample> samples;
    CerbIMULegPreint result;
};

static void mat3_mul(const double *A, const double *B, double *C) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j]; }
static void mat3_vec(const double *A, const double *v, double *o) { for (int i = 0; i < 3; i++) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2]; }
static void quat_to_R(const double *q, double *R) {      // (x, y, z, w), Eigen::Quaterniond::toRotationMatrix
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z, twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy; R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx; R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
static void R_to_quat(const double *m, double *q) {      // Eigen's Quaternion(Matrix3) constructor (vector2double: Quaterniond q{Rs[i]}, estimator.cpp:855)
    double t = m[0] + m[4] + m[8];
    if (t > 0) { t = std::sqrt(t + 1.0); q[3] = 0.5 * t; t = 0.5 / t; q[0] = (m[7] - m[5]) * t; q[1] = (m[2] - m[6]) * t; q[2] = (m[3] - m[1]) * t; }
    else {
        int i = 0; if (m[4] > m[0]) i = 1; if (m[8] > m[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        t = std::sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
        q[i] = 0.5 * t; t = 0.5 / t;
        q[3] = (m[3 * k + j] - m[3 * j + k]) * t; q[j] = (m[3 * j + i] + m[3 * i + j]) * t; q[k] = (m[3 * k + i] + m[3 * i + k]) * t;
    }
}

struct Robot {
    double Ps[NFRM][3], Vs[NFRM][3], Bas[NFRM][3], Bgs[NFRM][3], Rs[NFRM][9], Rho[NFRM][4], tic[2][3], ric[2][9], td = 0.0, g[3], Headers[NFRM];
    std::list<Feature> feature;
    int frame_count = 0, marginalization_flag = 0;
    Interval iv[NFRM];                     // il_pre_integrations[i]: frames i-1 -> i
    CerbIMULegSample last; bool has_last = false;
    bool has_prior = false; CerbPrior prior; std::vector<double> pJ, pr;
    bool openEx = false; int estimate_extrinsic = 1, estimate_td = 0;
    std::vector<double> path;              // per processed image: header, P(3), R(9), V(3), rho(4) of the newest frame
    std::vector<int> flag_hist;            // marginalization_flag of every processed image
    int last_track_num = 0, new_feature_num = 0, long_track_num = 0;

    Robot() {
        std::memset(Ps, 0, sizeof(Ps)); std::memset(Vs, 0, sizeof(Vs)); std::memset(Bas, 0, sizeof(Bas)); std::memset(Bgs, 0, sizeof(Bgs)); std::memset(tic, 0, sizeof(tic)); std::memset(Headers, 0, sizeof(Headers));
        for (int i = 0; i < NFRM; i++) { for (int k = 0; k < 9; k++) Rs[i][k] = (k % 4 == 0) ? 1.0 : 0.0; for (int k = 0; k < 4; k++) Rho[i][k] = 0.21; }
        for (int c = 0; c < 2; c++) for (int k = 0; k < 9; k++) ric[c][k] = (k % 4 == 0) ? 1.0 : 0.0;
        pJ.assign(CERB_MAX_PRIOR_DIM * CERB_MAX_PRIOR_DIM, 0.0); pr.assign(CERB_MAX_PRIOR_DIM, 0.0);
    }
    void new_interval(int j) { Interval &v = iv[j]; v = Interval(); v.valid = true; v.first = last; for (int k = 0; k < 3; k++) { v.ba[k] = Bas[j][k]; v.bg[k] = Bgs[j][k]; } for (int k = 0; k < 4; k++) v.rho[k] = Rho[j][k]; }

    // processIMULeg, estimator.cpp:590-653, one inter-frame interval at a time
    void process_interval(const CerbIMULegSample &first_sample, const CerbIMULegSample *smp, int n) {
        const int j = frame_count;
        if (!has_last) { last = first_sample; has_last = true; }
        if (!iv[j].valid) new_interval(j);
        if (j == 0) { if (n) last = smp[n - 1]; return; }
        Interval &v = iv[j];
        v.samples.insert(v.samples.end(), smp, smp + n); v.dirty = true;
        double acc_0[3], gyr_0[3], R[9], P[3], V[3];
        for (int k = 0; k < 3; k++) { acc_0[k] = last.acc[k]; gyr_0[k] = last.gyr[k]; P[k] = Ps[j][k]; V[k] = Vs[j][k]; }
        std::memcpy(R, Rs[j], sizeof(R));
        for (int s = 0; s < n; s++) {
            const double dt = smp[s].dt; const double *acc = smp[s].acc, *gyr = smp[s].gyr;
            double t[3], un_acc_0[3], un_gyr[3], un_acc_1[3];
            for (int k = 0; k < 3; k++) t[k] = acc_0[k] - Bas[j][k];
            mat3_vec(R, t, un_acc_0); for (int k = 0; k < 3; k++) un_acc_0[k] -= g[k];
            for (int k = 0; k < 3; k++) un_gyr[k] = 0.5 * (gyr_0[k] + gyr[k]) - Bgs[j][k];
            double dq[4] = {un_gyr[0] * dt / 2, un_gyr[1] * dt / 2, un_gyr[2] * dt / 2, 1.0}, dR[9], Rn[9];      // Utility::deltaQ, NOT normalised (as in the reference)
            quat_to_R(dq, dR); mat3_mul(R, dR, Rn); std::memcpy(R, Rn, sizeof(R));
            for (int k = 0; k < 3; k++) t[k] = acc[k] - Bas[j][k];
            mat3_vec(R, t, un_acc_1); for (int k = 0; k < 3; k++) un_acc_1[k] -= g[k];
            for (int k = 0; k < 3; k++) { const double un_acc = 0.5 * (un_acc_0[k] + un_acc_1[k]); P[k] = P[k] + dt * V[k] + 0.5 * dt * dt * un_acc; V[k] = V[k] + dt * un_acc; }
            for (int k = 0; k < 3; k++) { acc_0[k] = acc[k]; gyr_0[k] = gyr[k]; }
        }
        std::memcpy(Rs[j], R, sizeof(R)); for (int k = 0; k < 3; k++) { Ps[j][k] = P[k]; Vs[j][k] = V[k]; }
        if (n) last = smp[n - 1];
    }

    // FeatureManager::addFeatureCheckParallax, feature_manager.cpp:52-118 (image ordered by ascending feature id: the reference iterates a std::map)
    bool addFeatureCheckParallax(int fc, const CerbImage &im, double cur_td) {
        double parallax_sum = 0; int parallax_num = 0;
        last_track_num = 0; new_feature_num = 0; long_track_num = 0;
        std::map<int, Feature *> index; for (auto &it : feature) index[it.id] = &it;
        std::vector<int> order(im.n); for (int q = 0; q < im.n; q++) order[q] = q;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return im.ids[a] < im.ids[b]; });
        for (int q : order) {
            Obs f; const double *p = im.pts0 + 7 * q;
            f.point[0] = p[0]; f.point[1] = p[1]; f.point[2] = p[2]; f.velocity[0] = p[5]; f.velocity[1] = p[6]; f.cur_td = cur_td; f.is_stereo = false;
            f.pointRight[0] = f.pointRight[1] = f.pointRight[2] = 0; f.velocityRight[0] = f.velocityRight[1] = 0;
            if (im.has1[q]) { const double *r = im.pts1 + 7 * q; f.pointRight[0] = r[0]; f.pointRight[1] = r[1]; f.pointRight[2] = r[2]; f.velocityRight[0] = r[5]; f.velocityRight[1] = r[6]; f.is_stereo = true; }
            const int fid = (int)im.ids[q];
            auto it = index.find(fid);
            if (it == index.end()) { feature.emplace_back(); Feature &nf = feature.back(); nf.id = fid; nf.start_frame = fc; nf.obs.push_back(f); index[fid] = &nf; new_feature_num++; }
            else { it->second->obs.push_back(f); last_track_num++; if (it->second->obs.size() >= 4) long_track_num++; }
        }
        if (fc < 2 || last_track_num < 20 || long_track_num < 40 || new_feature_num > 0.5 * last_track_num) return true;
        for (auto &it : feature)
            if (it.start_frame <= fc - 2 && it.start_frame + (int)it.obs.size() - 1 >= fc - 1) {
                const Obs &fi = it.obs[fc - 2 - it.start_frame], &fj = it.obs[fc - 1 - it.start_frame];     // compensatedParallax2 (:531-565; the compensation is commented out in the reference)
                const double dep_i = fi.point[2], du = fi.point[0] / dep_i - fj.point[0], dv = fi.point[1] / dep_i - fj.point[1];
                parallax_sum += std::max(0.0, std::sqrt(std::min(du * du + dv * dv, du * du + dv * dv))); parallax_num++;
            }
        if (parallax_num == 0) return true;
        return parallax_sum / parallax_num >= kMinParallax;
    }
    void setDepth(const double *x) {                       // :142-160
        int k = -1;
        for (auto &it : feature) { it.used_num = (int)it.obs.size(); if (it.used_num < 4) continue; k++; it.estimated_depth = 1.0 / x[k]; it.solve_flag = it.estimated_depth < 0 ? 2 : 1; }
    }
    void removeFailures() { feature.remove_if([](const Feature &f) { return f.solve_flag == 2; }); }
    void removeBackShiftDepth(const std::map<int, double> &new_depth) {      // :450-488
        for (auto it = feature.begin(); it != feature.end();) {
            if (it->start_frame != 0) { it->start_frame--; ++it; continue; }
            it->obs.erase(it->obs.begin());
            if (it->obs.size() < 2) { it = feature.erase(it); continue; }
            it->estimated_depth = new_depth.at(it->id); ++it;
        }
    }
    void removeFront(int fc) {                             // :508-529
        for (auto it = feature.begin(); it != feature.end();) {
            if (it->start_frame == fc) { it->start_frame--; ++it; continue; }
            const int j = W - 1 - it->start_frame;
            if (it->endFrame() >= fc - 1) { it->obs.erase(it->obs.begin() + j); if (it->obs.empty()) { it = feature.erase(it); continue; } }
            ++it;
        }
    }
    void vector2double(CerbWindowState &st) const {        // estimator.cpp:848-901
        for (int i = 0; i < NFRM; i++) {
            for (int k = 0; k < 3; k++) { st.para_Pose[i][k] = Ps[i][k]; st.para_SpeedBias[i][k] = Vs[i][k]; st.para_SpeedBias[i][3 + k] = Bas[i][k]; st.para_SpeedBias[i][6 + k] = Bgs[i][k]; }
            R_to_quat(Rs[i], st.para_Pose[i] + 3);
            for (int k = 0; k < 4; k++) st.para_LegBias[i][k] = Rho[i][k];
        }
        for (int c = 0; c < 2; c++) { for (int k = 0; k < 3; k++) st.para_Ex_Pose[c][k] = tic[c][k]; R_to_quat(ric[c], st.para_Ex_Pose[c] + 3); }
        st.para_Td[0] = td;
    }
    int depthVector(double *out) {                         // getDepthVector :180-196
        int k = 0; for (auto &it : feature) { it.used_num = (int)it.obs.size(); if (it.used_num < 4) continue; out[k++] = 1.0 / it.estimated_depth; }
        return k;
    }
    void double2vector_rest(const CerbWindowState &st, const double *para_feature, const double *P, const double *R, const double *V) {     // :936-1003 after the gauge fix
        std::memcpy(Ps, P, sizeof(Ps)); std::memcpy(Rs, R, sizeof(Rs)); std::memcpy(Vs, V, sizeof(Vs));
        for (int i = 0; i < NFRM; i++) { for (int k = 0; k < 3; k++) { Bas[i][k] = st.para_SpeedBias[i][3 + k]; Bgs[i][k] = st.para_SpeedBias[i][6 + k]; } for (int k = 0; k < 4; k++) Rho[i][k] = st.para_LegBias[i][k]; }
        for (int c = 0; c < 2; c++) {
            for (int k = 0; k < 3; k++) tic[c][k] = st.para_Ex_Pose[c][k];
            double q[4]; const double *s = st.para_Ex_Pose[c] + 3; const double nrm = std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2] + s[3] * s[3]);
            for (int k = 0; k < 4; k++) q[k] = s[k] / nrm;
            quat_to_R(q, ric[c]);
        }
        setDepth(para_feature);
        td = st.para_Td[0];
    }
    void slide_window(const std::map<int, double> &new_depth) {             // estimator.cpp:1460-1677 (frame_count == WINDOW_SIZE, USE_LEG && USE_IMU)
        if (marginalization_flag == 0) {
            for (int i = 0; i < W; i++) {
                Headers[i] = Headers[i + 1]; std::memcpy(Rs[i], Rs[i + 1], sizeof(Rs[i])); std::memcpy(Ps[i], Ps[i + 1], sizeof(Ps[i])); std::memcpy(Vs[i], Vs[i + 1], sizeof(Vs[i]));
                std::memcpy(Bas[i], Bas[i + 1], sizeof(Bas[i])); std::memcpy(Bgs[i], Bgs[i + 1], sizeof(Bgs[i])); std::memcpy(Rho[i], Rho[i + 1], sizeof(Rho[i]));
            }
            for (int i = 0; i < W; i++) iv[i] = std::move(iv[i + 1]);
            new_interval(W);
            removeBackShiftDepth(new_depth);
        } else {
            Headers[W - 1] = Headers[W]; std::memcpy(Ps[W - 1], Ps[W], sizeof(Ps[W])); std::memcpy(Rs[W - 1], Rs[W], sizeof(Rs[W]));
            Interval &a = iv[W - 1], &b = iv[W];
            a.samples.insert(a.samples.end(), b.samples.begin(), b.samples.end()); a.dirty = true;
            std::memcpy(Vs[W - 1], Vs[W], sizeof(Vs[W])); std::memcpy(Bas[W - 1], Bas[W], sizeof(Bas[W])); std::memcpy(Bgs[W - 1], Bgs[W], sizeof(Bgs[W])); std::memcpy(Rho[W - 1], Rho[W], sizeof(Rho[W]));
            new_interval(W);
            removeFront(frame_count);
        }
    }
};

}  // namespace cerbhost

struct CerbReplay {
    CerbHandle *h; CerbPreintConfig pcfg; int n, F;
    std::vector<cerbhost::Robot> robots;
    // host batches (solve: features with >= 4 observations; all: every track, for triangulation and the depth shift)
    std::vector<CerbWindowDesc> descs, descs_all; std::vector<CerbWindowState> states, states_all, before;
    std::vector<CerbFeature> feats, feats_all; std::vector<CerbObservation> obs, obs_all; std::vector<double> lam, lam_all;
    std::vector<CerbIMULegPreint> preint; std::vector<int> ids, ids_all, nids, nids_all;
    std::vector<CerbPrior> next_priors; std::vector<double> nJ, nr;
    double t_device[6] = {0, 0, 0, 0, 0, 0}, t_host = 0;       // preintegrate, triangulate, solve, marginalize, outliers, shift
};

namespace cerbhost {
// the robots are independent: per-robot bookkeeping runs on a few host threads (fn returns a status; the first failure wins)
template <class Fn> static int parallel_robots(int n, Fn fn) {
    unsigned hw = std::thread::hardware_concurrency();
    int nth = (int)std::min<unsigned>(hw ? hw : 1, 16u);
    if (n < 16) nth = 1;
    std::vector<int> rcs(nth, CERB_OK); std::vector<std::string> errs(nth);
    auto work = [&](int t) { for (int w = t; w < n; w += nth) { const int rc = fn(w); if (rc) { rcs[t] = rc; errs[t] = g_err; return; } } };
    std::vector<std::thread> th;
    for (int t = 1; t < nth; t++) th.emplace_back(work, t);
    work(0);
    for (auto &t : th) t.join();
    for (int t = 0; t < nth; t++) if (rcs[t]) return fail(rcs[t], errs[t]);
    return CERB_OK;
}
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// CerbWindowDesc / CerbWindowState of robot w: the factor enumeration of estimator.cpp:1114-1216 (features with used_num >= min_used in list order)
static int fill_window(CerbReplay *rp, int w, int min_used, bool all) {
    Robot &e = rp->robots[w];
    const int Fcap = all ? 2 * rp->F : rp->F, Ocap = Fcap * NFRM;
    CerbFeature *fw = (all ? rp->feats_all.data() : rp->feats.data()) + (size_t)w * Fcap;
    CerbObservation *ow = (all ? rp->obs_all.data() : rp->obs.data()) + (size_t)w * Ocap;
    int *idw = (all ? rp->ids_all.data() : rp->ids.data()) + (size_t)w * Fcap;
    int k = 0, off = 0;
    for (auto &it : e.feature) {
        it.used_num = (int)it.obs.size();
        if (it.used_num < min_used) continue;
        if (k >= Fcap || off + it.used_num > Ocap) return fail(CERB_ERR_BAD_ARGUMENT, "replay: window over the batch capacity");
        fw[k].start_frame = it.start_frame; fw[k].n_obs = it.used_num; fw[k].obs_offset = off; fw[k].reserved = 0;
        for (const Obs &f : it.obs) {
            CerbObservation &o = ow[off++];
            o.point[0] = f.point[0]; o.point[1] = f.point[1]; o.velocity[0] = f.velocity[0]; o.velocity[1] = f.velocity[1];
            o.pointRight[0] = f.pointRight[0]; o.pointRight[1] = f.pointRight[1]; o.velocityRight[0] = f.velocityRight[0]; o.velocityRight[1] = f.velocityRight[1];
            o.cur_td = f.cur_td; o.is_stereo = f.is_stereo ? 1 : 0; o.reserved = 0;
        }
        idw[k++] = it.id;
    }
    (all ? rp->nids_all : rp->nids)[w] = k;
    CerbWindowDesc &d = (all ? rp->descs_all : rp->descs)[w];
    std::memset(&d, 0, sizeof(d));
    d.n_features = k; d.n_obs = off; d.features = fw; d.obs = ow;
    CerbIMULegPreint *pw = rp->preint.data() + (size_t)w * W;
    for (int i = 0; i < W; i++) pw[i] = e.iv[i + 1].result;
    d.preint = pw;
    CerbWindowState &st = (all ? rp->states_all : rp->states)[w];
    e.vector2double(st);
    st.para_Feature = (all ? rp->lam_all.data() : rp->lam.data()) + (size_t)w * Fcap;
    if (!all) {
        e.depthVector(st.para_Feature);
        const double v0 = std::sqrt(e.Vs[0][0] * e.Vs[0][0] + e.Vs[0][1] * e.Vs[0][1] + e.Vs[0][2] * e.Vs[0][2]);
        if (e.estimate_extrinsic && e.frame_count == W && v0 > 0.2) e.openEx = true;                  // estimator.cpp:1091-1100
        d.extrinsic_open = (e.estimate_extrinsic && e.openEx) ? 1 : 0;
        d.td_open = (e.estimate_td && v0 >= 0.2) ? 1 : 0;
        if (e.has_prior && e.prior.valid) { d.prior = e.prior; d.prior.linearized_jacobians = e.pJ.data(); d.prior.linearized_residuals = e.pr.data(); }
    }
    return CERB_OK;
}

static int preintegrate_dirty(CerbReplay *rp) {
    std::vector<CerbPreintJob> jobs; std::vector<std::pair<int, int>> who;
    for (int w = 0; w < rp->n; w++) for (int i = 1; i <= W; i++) {
        Interval &v = rp->robots[w].iv[i];
        if (!v.valid || !v.dirty || v.samples.empty()) continue;
        CerbPreintJob j; std::memset(&j, 0, sizeof(j));
        std::memcpy(j.acc_0, v.first.acc, 24); std::memcpy(j.gyr_0, v.first.gyr, 24); std::memcpy(j.phi_0, v.first.phi, 96); std::memcpy(j.dphi_0, v.first.dphi, 96); std::memcpy(j.c_0, v.first.c, 32);
        std::memcpy(j.linearized_ba, v.ba, 24); std::memcpy(j.linearized_bg, v.bg, 24); std::memcpy(j.linearized_rho, v.rho, 32);
        j.n_samples = (int)v.samples.size(); j.samples = v.samples.data();
        jobs.push_back(j); who.emplace_back(w, i);
    }
    if (jobs.empty()) return CERB_OK;
    std::vector<CerbIMULegPreint> out(jobs.size());
    const double t0 = now_s();
    int rc = cerb_preintegrate_batch(rp->h, &rp->pcfg, (int)jobs.size(), jobs.data(), out.data()); if (rc) return rc;
    rp->t_device[0] += now_s() - t0;
    for (size_t q = 0; q < jobs.size(); q++) { Interval &v = rp->robots[who[q].first].iv[who[q].second]; v.result = out[q]; v.has_result = true; v.dirty = false; }
    return CERB_OK;
}
}  // namespace cerbhost

extern "C" {

int cerb_replay_create(CerbHandle *h, const CerbPreintConfig *pcfg, int32_t n_robots, int32_t max_features, int32_t estimate_extrinsic, int32_t estimate_td, CerbReplay **out) {
    if (!h || !pcfg || !out || n_robots < 1 || max_features < 1) return fail(CERB_ERR_BAD_ARGUMENT, "cerb_replay_create: bad argument");
    if (n_robots > h->B || 2 * max_features > h->F) return fail(CERB_ERR_BAD_ARGUMENT, "cerb_replay_create: the handle needs max_batch >= n_robots and max_features >= 2 x the replay's (the triangulation batch holds every track)");
    CerbReplay *rp = new CerbReplay();
    rp->h = h; rp->pcfg = *pcfg; rp->n = n_robots; rp->F = max_features;
    rp->robots.resize(n_robots);
    for (auto &e : rp->robots) { for (int k = 0; k < 3; k++) e.g[k] = h->cfg.g[k]; e.estimate_extrinsic = estimate_extrinsic; e.estimate_td = estimate_td; }
    const size_t n = n_robots, F = max_features;
    rp->descs.resize(n); rp->descs_all.resize(n); rp->states.resize(n); rp->states_all.resize(n); rp->before.resize(n);
    rp->feats.resize(n * F); rp->feats_all.resize(n * 2 * F); rp->obs.resize(n * F * cerbhost::NFRM); rp->obs_all.resize(n * 2 * F * cerbhost::NFRM);
    rp->lam.assign(n * F, 0.0); rp->lam_all.assign(n * 2 * F, 0.0); rp->preint.resize(n * cerbhost::W);
    rp->ids.assign(n * F, 0); rp->ids_all.assign(n * 2 * F, 0); rp->nids.assign(n, 0); rp->nids_all.assign(n, 0);
    rp->next_priors.resize(n); rp->nJ.assign(n * CERB_MAX_PRIOR_DIM * CERB_MAX_PRIOR_DIM, 0.0); rp->nr.assign(n * CERB_MAX_PRIOR_DIM, 0.0);
    *out = rp;
    return CERB_OK;
}
void cerb_replay_destroy(CerbReplay *rp) { delete rp; }

int cerb_replay_set_extrinsics(CerbReplay *rp, int32_t robot, const double *tic, const double *ric) {
    if (!rp || robot < 0 || robot >= rp->n || !tic || !ric) return fail(CERB_ERR_BAD_ARGUMENT, "cerb_replay_set_extrinsics: bad argument");
    std::memcpy(rp->robots[robot].tic, tic, 6 * sizeof(double)); std::memcpy(rp->robots[robot].ric, ric, 18 * sizeof(double));
    return CERB_OK;
}

int cerb_replay_seed_frame(CerbReplay *rp, int32_t robot, int32_t k, const double *P, const double *R, const double *V, const CerbIMULegSample *first,
                           const CerbIMULegSample *samples, int32_t n_samples, const CerbImage *image, double header) {
    if (!rp || robot < 0 || robot >= rp->n || k < 0 || k > cerbhost::W || !P || !R || !V || !first) return fail(CERB_ERR_BAD_ARGUMENT, "cerb_replay_seed_frame: bad argument");
    cerbhost::Robot &e = rp->robots[robot];
    e.frame_count = k;
    std::memcpy(e.Ps[k], P, 24); std::memcpy(e.Rs[k], R, 72); std::memcpy(e.Vs[k], V, 24);
    e.process_interval(*first, samples, k == 0 ? 0 : n_samples);
    std::memcpy(e.Ps[k], P, 24); std::memcpy(e.Rs[k], R, 72); std::memcpy(e.Vs[k], V, 24);          // seeded states, not the IMU prediction
    e.Headers[k] = header;
    if (k < cerbhost::W && image) e.addFeatureCheckParallax(k, *image, e.td);
    if (k == cerbhost::W) e.frame_count = cerbhost::W;
    return CERB_OK;
}

// processMeasurements for one camera frame of every robot (NON_LINEAR): images[w], firsts[w] (sample at the previous frame instant), samples[w]
int cerb_replay_step(CerbReplay *rp, const CerbImage *images, const CerbIMULegSample *firsts, const CerbIMULegSample *const *samples, const int32_t *n_samples, double header,
                     CerbSolveReport *reports) {
    using namespace cerbhost;
    if (!rp || !images || !firsts || !samples || !n_samples) return fail(CERB_ERR_BAD_ARGUMENT, "cerb_replay_step: null argument");
    const int n = rp->n, F = rp->F;
    int rc;
    double th = now_s(), t0;
    parallel_robots(n, [&](int w) {
        Robot &e = rp->robots[w];
        e.process_interval(firsts[w], samples[w], n_samples[w]);
        e.Headers[e.frame_count] = header;
        e.marginalization_flag = e.addFeatureCheckParallax(e.frame_count, images[w], e.td) ? 0 : 1;
        e.flag_hist.push_back(e.marginalization_flag);
        return (int)CERB_OK;
    });
    rp->t_host += now_s() - th;
    rc = preintegrate_dirty(rp); if (rc) return rc;
    // ---- f_manager.triangulate (estimator.cpp:803)
    th = now_s();
    rc = parallel_robots(n, [&](int w) {
        const int r2 = fill_window(rp, w, 1, true); if (r2) return r2;
        double *lam = rp->lam_all.data() + (size_t)w * 2 * F; int k = 0;
        for (auto &it : rp->robots[w].feature) lam[k++] = it.estimated_depth > 0 ? 1.0 / it.estimated_depth : -1.0;
        return (int)CERB_OK;
    });
    if (rc) return rc;
    rp->t_host += now_s() - th; t0 = now_s();
    std::vector<double> depth((size_t)n * rp->h->F, 0.0);
    rc = cerb_batch_upload(rp->h, n, rp->descs_all.data(), rp->states_all.data()); if (rc) return rc;
    rc = cerb_batch_triangulate(rp->h, kInitDepth, depth.data()); if (rc) return rc;
    rp->t_device[1] += now_s() - t0; th = now_s();
    rc = parallel_robots(n, [&](int w) {
        int k = 0; for (auto &it : rp->robots[w].feature) { if (!(it.estimated_depth > 0)) it.estimated_depth = depth[(size_t)w * rp->h->F + k]; k++; }
        return fill_window(rp, w, 4, false);           // ---- optimization(): solve
    });
    if (rc) return rc;
    std::memcpy(rp->before.data(), rp->states.data(), sizeof(CerbWindowState) * n);
    rp->t_host += now_s() - th; t0 = now_s();
    std::vector<CerbSolveReport> rep(n);
    rc = cerb_solve_batch(rp->h, n, rp->descs.data(), rp->states.data(), rep.data()); if (rc && rc != CERB_ERR_NON_FINITE) return rc;
    rp->t_device[2] += now_s() - t0; th = now_s();
    if (reports) std::memcpy(reports, rep.data(), sizeof(CerbSolveReport) * n);
    // ---- optimization(): marginalization at the re-anchored states (vector2double again, estimator.cpp:1251 / :1384); the batch is still resident
    std::vector<int32_t> flags(n);
    parallel_robots(n, [&](int w) {
        double P[NFRM * 3], R[NFRM * 9], V[NFRM * 3];
        cerb_double2vector(&rp->before[w], &rp->states[w], P, R, V);
        rp->robots[w].double2vector_rest(rp->states[w], rp->states[w].para_Feature, P, R, V);
        Robot &e = rp->robots[w];
        double *keep = rp->states[w].para_Feature;
        e.vector2double(rp->states[w]); rp->states[w].para_Feature = keep; e.depthVector(keep);
        flags[w] = e.marginalization_flag;
        rp->next_priors[w].linearized_jacobians = rp->nJ.data() + (size_t)w * CERB_MAX_PRIOR_DIM * CERB_MAX_PRIOR_DIM; rp->next_priors[w].linearized_residuals = rp->nr.data() + (size_t)w * CERB_MAX_PRIOR_DIM;
        return (int)CERB_OK;
    });
    rp->t_host += now_s() - th; t0 = now_s();
    rc = cerb_batch_marginalize(rp->h, flags.data(), rp->states.data(), rp->next_priors.data(), nullptr); if (rc) return rc;
    rp->t_device[3] += now_s() - t0; th = now_s();
    for (int w = 0; w < n; w++) {
        Robot &e = rp->robots[w]; const CerbPrior &np = rp->next_priors[w];
        if (np.valid) {
            e.prior = np; std::memcpy(e.pJ.data(), np.linearized_jacobians, sizeof(double) * np.n * np.n); std::memcpy(e.pr.data(), np.linearized_residuals, sizeof(double) * np.n);
            e.has_prior = true;
        } else if (flags[w] == 0) e.has_prior = false;       // MARGIN_SECOND_NEW without a prior keeps "none"; MARGIN_OLD with nothing dropped: valid = false
    }
    // ---- outliersRejection + removeOutlier (:812-814) at the re-anchored states
    rp->t_host += now_s() - th; t0 = now_s();
    std::vector<double> err((size_t)n * rp->h->F, 0.0);
    rc = cerb_batch_update_states(rp->h, n, rp->states.data()); if (rc) return rc;       // same windows as the solve, moved by double2vector: only the states travel
    rc = cerb_batch_outlier_errors(rp->h, kFocal, err.data(), nullptr); if (rc) return rc;
    rp->t_device[4] += now_s() - t0; th = now_s();
    rc = parallel_robots(n, [&](int w) {
        Robot &e = rp->robots[w]; const int *idw = rp->ids.data() + (size_t)w * F;
        std::map<int, bool> bad; bool any = false;
        for (int k = 0; k < rp->nids[w]; k++) if (err[(size_t)w * rp->h->F + k] * kFocal > 3) { bad[idw[k]] = true; any = true; }
        if (any) e.feature.remove_if([&](const Feature &f) { return bad.count(f.id) != 0; });
        // ---- slideWindow (+ removeBackShiftDepth on the device for the robots that marginalize the oldest frame)
        const int r2 = fill_window(rp, w, 1, true); if (r2) return r2;
        double *lam = rp->lam_all.data() + (size_t)w * 2 * F; int k = 0;
        for (auto &it : e.feature) lam[k++] = 1.0 / it.estimated_depth;
        return (int)CERB_OK;
    });
    if (rc) return rc;
    rp->t_host += now_s() - th; t0 = now_s();
    std::vector<int32_t> nstart((size_t)n * rp->h->F), keepf((size_t)n * rp->h->F); std::vector<double> sdepth((size_t)n * rp->h->F, 0.0);
    rc = cerb_batch_upload(rp->h, n, rp->descs_all.data(), rp->states_all.data()); if (rc) return rc;
    rc = cerb_batch_shift_depth(rp->h, kInitDepth, nstart.data(), sdepth.data(), keepf.data()); if (rc) return rc;
    rp->t_device[5] += now_s() - t0; th = now_s();
    parallel_robots(n, [&](int w) {
        Robot &e = rp->robots[w]; const int *idw = rp->ids_all.data() + (size_t)w * 2 * F;
        std::map<int, double> nd; for (int k = 0; k < rp->nids_all[w]; k++) nd[idw[k]] = sdepth[(size_t)w * rp->h->F + k];
        e.slide_window(nd);
        e.removeFailures();
        e.path.push_back(header); for (int k = 0; k < 3; k++) e.path.push_back(e.Ps[W][k]); for (int k = 0; k < 9; k++) e.path.push_back(e.Rs[W][k]);
        for (int k = 0; k < 3; k++) e.path.push_back(e.Vs[W][k]); for (int k = 0; k < 4; k++) e.path.push_back(e.Rho[W][k]);
        return (int)CERB_OK;
    });
    rp->t_host += now_s() - th;
    return CERB_OK;
}

int cerb_replay_path(CerbReplay *rp, int32_t robot, int32_t *n_rows, double *out, int32_t max_rows) {
    if (!rp || robot < 0 || robot >= rp->n || !n_rows) return fail(CERB_ERR_BAD_ARGUMENT, "cerb_replay_path: bad argument");
    const std::vector<double> &p = rp->robots[robot].path;
    *n_rows = (int32_t)(p.size() / 20);
    if (out) { const size_t rows = std::min<size_t>(*n_rows, max_rows > 0 ? max_rows : 0); std::memcpy(out, p.data(), rows * 20 * sizeof(double)); }
    return CERB_OK;
}
int cerb_replay_flags(CerbReplay *rp, int32_t robot, int32_t *n, int32_t *flags, int32_t max_flags) {
    if (!rp || robot < 0 || robot >= rp->n || !n) return fail(CERB_ERR_BAD_ARGUMENT, "cerb_replay_flags: bad argument");
    const std::vector<int> &f = rp->robots[robot].flag_hist;
    *n = (int32_t)f.size();
    if (flags) for (int k = 0; k < *n && k < max_flags; k++) flags[k] = f[k];
    return CERB_OK;
}
int cerb_replay_timing(CerbReplay *rp, double *device6, double *host) {
    if (!rp) return fail(CERB_ERR_BAD_ARGUMENT, "null replay");
    if (device6) std::memcpy(device6, rp->t_device, sizeof(rp->t_device));
    if (host) *host = rp->t_host;
    return CERB_OK;
}
int cerb_replay_feature_ids(CerbReplay *rp, int32_t robot, int32_t *n, int32_t *ids, int32_t max_ids) {
    if (!rp || robot < 0 || robot >= rp->n || !n) return fail(CERB_ERR_BAD_ARGUMENT, "cerb_replay_feature_ids: bad argument");
    int k = 0; for (auto &it : rp->robots[robot].feature) { if (ids && k < max_ids) ids[k] = it.id; k++; }
    *n = k;
    return CERB_OK;
}

}  // extern "C"
