// TEST INFRASTRUCTURE ONLY -- single-threaded host build of smplify-x-partial_b200/csrc/
// sfx_core.cuh, used by tests/test_hostsim_*.py to debug the evaluation maths and the
// optimiser control flow without a GPU.  It is never linked into, or reachable from, the
// product library (libsfx.so has no CPU path: it fails when CUDA is unavailable).
#include <cstring>
#include <memory>
#include <string>
#include <vector>

static std::vector<double> g_trace;
#define SFX_TRACE(v) g_trace.push_back(v)
static std::vector<double> g_steps;
static int g_hits_cap = 0;      // 0: the device's region size
static int g_all_rows = 1;      // 1: a single evaluation streams every support row (it returns all joints,
                                // like the device's eval with a joints output); 0: only the live rows
#define SFX_TRACE_STEP(t) g_steps.push_back(t)
#include "../../smplify-x-partial_b200/csrc/sfx_model_prep.h"

using namespace sfx;

template <typename T>
struct SimModel {
    HostModel<T> h;
    int gmm_M = 0, gmm_D = 0;
    std::vector<T> gmm_means, gmm_prec, gmm_logw;
    std::vector<T> vw1, vb1, vw2, vb2, vw3, vb3;
    HostCollision coll;
    bool has_coll = false;
    int coll_work_bytes = 131072;
    int last_touch = 0, last_overflow = 0;
};
struct SimHandle {
    int use_double;
    SimModel<float> f;
    SimModel<double> d;
};

template <typename T>
static void run(SimModel<T>& sm, const SfxStage* st, int use_vposer, int do_fit, T* params,
                const T* gt, const T* conf, const T* jw, const unsigned char* lowconf,
                const unsigned char* init_mask, const T* cam, const T* reg_pose, T* loss_out,
                T* grad_out, T* joints_out, int* n_evals, int* flags) {
    ModelView<T> M = sm.h.host_view();
    M.gmm_M = sm.gmm_M; M.gmm_D = sm.gmm_D;
    if (!sm.vw1.empty()) {
        M.vp_ready = 1; M.vp_w1 = sm.vw1.data(); M.vp_b1 = sm.vb1.data(); M.vp_w2 = sm.vw2.data();
        M.vp_b2 = sm.vb2.data(); M.vp_w3 = sm.vw3.data(); M.vp_b3 = sm.vb3.data();
    }
    M.gmm_means = sm.gmm_means.data(); M.gmm_prec = sm.gmm_prec.data(); M.gmm_logw = sm.gmm_logw.data();
    SfxLayout L = make_layout(sm.h.NB, sm.h.NE, sm.h.NH, use_vposer);
    std::unique_ptr<Scratch<T>> Sp(new Scratch<T>());
    Scratch<T>& S = *Sp;
    std::memset(&S, 0, sizeof(S));
    for (int i = 0; i < L.np; ++i) S.x[i] = params[i];
    support_begin_frame(M, S);
    stage_setup(M, *st, jw, lowconf, conf, init_mask, M.K, S, joints_out != nullptr && !do_fit && g_all_rows);
    std::vector<T> hs((size_t)SFX_HIST * SFX_NP_MAX), hy((size_t)SFX_HIST * SFX_NP_MAX);
    CollWS<T> W;
    std::vector<T> vp_g, vert_g, dvert_g, dtri_g, fbox_buf, ftri;
    std::vector<unsigned short> tv_g, hits_g;
    std::vector<unsigned char> work, sort_g;
    const CollWS<T>* Wp = nullptr;
    if (sm.has_coll) {
        M.coll_ready = 1; M.n_parts = sm.coll.n_parts; M.faces = sm.h.faces.data();
        M.part_ptr = sm.coll.part_ptr.data(); M.part_faces = sm.coll.part_faces.data();
        M.face_part = sm.coll.face_part.data(); M.part_allow = sm.coll.part_allow.data();
        M.vf_ptr = sm.coll.vf_ptr.data(); M.vf_idx = sm.coll.vf_idx.data();
        M.n_clusters = sm.coll.n_clusters; M.cl_ptr = sm.coll.cl_ptr.data();
        M.part_cl_ptr = sm.coll.part_cl_ptr.data();
        vp_g.assign((size_t)3 * M.V, 0); vert_g.assign((size_t)3 * M.V, 0);
        dvert_g.assign((size_t)3 * M.V, 0); dtri_g.assign((size_t)9 * M.F, 0);
        work.assign(sm.coll_work_bytes, 0);
        fbox_buf.assign((size_t)6 * M.F, 0); sort_g.assign((size_t)SFX_COLL_ENTRY * SFX_COLL_SORT_G, 0); tv_g.assign(M.V, 0);
        ftri.assign((size_t)9 * M.F, 0);
        W.fbox = fbox_buf.data(); W.ftri = ftri.data(); W.sort_g = sort_g.data(); W.tv_g = tv_g.data();
        W.hits_cap = g_hits_cap > 0 ? g_hits_cap : SFX_COLL_HITS;
        hits_g.assign((size_t)W.hits_cap, 0);         // one thread on the host: many chunks
        W.hits_g = hits_g.data();
        W.vp_g = vp_g.data(); W.vert_g = vert_g.data(); W.dvert_g = dvert_g.data();
        W.dtri_g = dtri_g.data(); W.work = work.data(); W.work_bytes = sm.coll_work_bytes;
        Wp = &W;
    }
    std::vector<T> gram((size_t)2 * SFX_HIST * SFX_HIST);
    EvalCtx<T> E{&M, &L, st, gt, conf, init_mask, cam, reg_pose, nullptr, Wp, gram.data()};
    if (!do_fit) {
        eval_frame(M, L, *st, gt, conf, init_mask, cam, reg_pose, S, nullptr, Wp);
        *loss_out = S.loss;
        for (int i = 0; i < L.np; ++i) grad_out[i] = S.gfull[i];
    } else {
        double r = run_fitting(E, S, hs.data(), hy.data(), flags);
        *loss_out = (T)r;
        for (int i = 0; i < L.np; ++i) params[i] = S.x[i];
    }
    if (joints_out)
        for (int k = 0; k < M.K; ++k)
            for (int a = 0; a < 3; ++a) joints_out[3 * k + a] = S.X[3 * M.joint_map[k] + a];
    *n_evals = S.n_evals;
    sm.last_touch = S.n_touch;
    sm.last_overflow = S.coll_overflow;
    if (S.coll_overflow) *flags |= SFX_FLAG_COLL_OVERFLOW;
}

template <typename T>
static void set_gmm(SimModel<T>& sm, int M, int D, const T* means, const T* prec, const T* logw) {
    sm.gmm_M = M; sm.gmm_D = D;
    sm.gmm_means.assign(means, means + (size_t)M * D);
    sm.gmm_prec.assign(prec, prec + (size_t)M * D * D);
    sm.gmm_logw.assign(logw, logw + M);
}
template <typename T>
static void set_vposer(SimModel<T>& sm, const T* w1, const T* b1, const T* w2, const T* b2,
                       const T* w3, const T* b3) {
    sm.vw1.assign(w1, w1 + 512 * 32); sm.vb1.assign(b1, b1 + 512);
    sm.vw2.assign(w2, w2 + 512 * 512); sm.vb2.assign(b2, b2 + 512);
    sm.vw3.assign((size_t)128 * 512, (T)0); sm.vb3.assign(128, (T)0);
    for (size_t i = 0; i < (size_t)126 * 512; ++i) sm.vw3[i] = w3[i];
    for (int i = 0; i < 126; ++i) sm.vb3[i] = b3[i];
}
// Unit access to the L-BFGS direction (lbfgs_ls.py:336-358): k pairs (s, y) [k][D], oldest first,
// gradient g, initial scaling hd.  mode 0: the reference's recursion, statement for statement
// the block-wide path of lbfgs_step; mode 2: gram_two_loop, its Gram blocks built pair by pair
// exactly as lbfgs_step builds them (one call per stored pair).
template <typename T>
static void lbfgs_direction(int mode, int k, int D, const T* s, const T* y, const T* g, T hd, T* d_out) {
    std::unique_ptr<Scratch<T>> Sp(new Scratch<T>());
    Scratch<T>& S = *Sp;
    std::memset(&S, 0, sizeof(S));
    const int H = SFX_HIST;
    std::vector<T> hs((size_t)SFX_HIST * SFX_NP_MAX), hy((size_t)SFX_HIST * SFX_NP_MAX);
    std::vector<T> gram((size_t)2 * SFX_HIST * SFX_HIST);
    for (int i = 0; i < k; ++i) {
        for (int e = 0; e < D; ++e) {
            hs[(size_t)i * SFX_NP_MAX + e] = s[(size_t)i * D + e];
            hy[(size_t)i * SFX_NP_MAX + e] = y[(size_t)i * D + e];
        }
        T ys = block_dot(&hy[(size_t)i * SFX_NP_MAX], &hs[(size_t)i * SFX_NP_MAX], D, &S.red[2]);
        S.ro[i] = (T)1 / ys;
    }
    for (int e = 0; e < D; ++e) S.g[e] = g[e];
    if (mode == 0) {
        for (int e = 0; e < D; ++e) S.q[e] = -S.g[e];
        for (int i = k - 1; i >= 0; --i) {
            const T* srow = &hs[(size_t)i * SFX_NP_MAX];
            const T* yrow = &hy[(size_t)i * SFX_NP_MAX];
            T a = block_dot(srow, S.q, D, &S.red[2]) * S.ro[i];
            S.al[i] = a;
            for (int e = 0; e < D; ++e) S.q[e] += -a * yrow[e];
        }
        for (int e = 0; e < D; ++e) S.d[e] = S.q[e] * hd;
        for (int i = 0; i < k; ++i) {
            const T* srow = &hs[(size_t)i * SFX_NP_MAX];
            const T* yrow = &hy[(size_t)i * SFX_NP_MAX];
            T be = block_dot(yrow, S.d, D, &S.red[2]) * S.ro[i];
            T co = S.al[i] - be;
            for (int e = 0; e < D; ++e) S.d[e] += co * srow[e];
        }
    } else {
        EvalCtx<T> E{};
        E.gram = gram.data();
        for (int i = 0; i < k; ++i) {
            for (int e = 0; e < D; ++e) {
                S.x0[e] = hs[(size_t)i * SFX_NP_MAX + e];
                S.q[e] = hy[(size_t)i * SFX_NP_MAX + e];
            }
            gram_two_loop(E, S, i + 1, 0, H, hd, hs.data(), hy.data(), D, true);
        }
        if (k == 0) gram_two_loop(E, S, 0, 0, H, hd, hs.data(), hy.data(), D, false);
    }
    for (int e = 0; e < D; ++e) d_out[e] = S.d[e];
}

extern "C" {
void hs_lbfgs_direction(int use_double, int mode, int k, int D, const void* s, const void* y,
                        const void* g, double hd, void* d_out) {
    if (use_double)
        lbfgs_direction<double>(mode, k, D, (const double*)s, (const double*)y, (const double*)g, hd, (double*)d_out);
    else
        lbfgs_direction<float>(mode, k, D, (const float*)s, (const float*)y, (const float*)g, (float)hd, (float*)d_out);
}
void hs_set_vposer(void* p, const void* w1, const void* b1, const void* w2, const void* b2,
                   const void* w3, const void* b3) {
    SimHandle* h = (SimHandle*)p;
    if (h->use_double)
        set_vposer<double>(h->d, (const double*)w1, (const double*)b1, (const double*)w2,
                           (const double*)b2, (const double*)w3, (const double*)b3);
    else
        set_vposer<float>(h->f, (const float*)w1, (const float*)b1, (const float*)w2,
                          (const float*)b2, (const float*)w3, (const float*)b3);
}
// face segmentation + ignored part pairs of the interpenetration term; work_bytes = size of the
// candidate area (the device uses the idle blend ring: 128 KB in float, 64 KB in double);
// parents == NULL: the term without FilterFaces (segm only groups the faces)
int hs_set_collision(void* p, const int32_t* segm, const int32_t* parents, const int32_t* ign,
                     int n_ign, int work_bytes, char* err, int errlen) {
    SimHandle* h = (SimHandle*)p;
    std::string e;
    if (h->use_double) {
        e = prepare_collision(h->d.h.V, h->d.h.F, h->d.h.faces.data(), h->d.h.vt.data(), segm, parents, ign, n_ign, h->d.coll, parents == nullptr);
        h->d.has_coll = e.empty(); h->d.coll_work_bytes = work_bytes;
    } else {
        e = prepare_collision(h->f.h.V, h->f.h.F, h->f.h.faces.data(), h->f.h.vt.data(), segm, parents, ign, n_ign, h->f.coll, parents == nullptr);
        h->f.has_coll = e.empty(); h->f.coll_work_bytes = work_bytes;
    }
    if (!e.empty()) { std::strncpy(err, e.c_str(), errlen - 1); return -1; }
    return 0;
}
// unit access to the narrow phase and the pair penalty (double)
int hs_triangles_intersect(const double* t1, const double* t2) { return triangles_intersect(t1, t2) ? 1 : 0; }
void hs_pair_terms(const double* ti, const double* tj, double sigma, double* loss, double* gi) {
    pair_terms(ti, tj, sigma, loss, gi);
}
void hs_set_hits_cap(int n) { g_hits_cap = n; }
void hs_set_all_rows(int on) { g_all_rows = on; }
int hs_last_touch(void* p) { SimHandle* h = (SimHandle*)p; return h->use_double ? h->d.last_touch : h->f.last_touch; }
int hs_trace(double* out, int cap) {
    int n = (int)g_trace.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = g_trace[i];
    g_trace.clear();
    return n;
}

int hs_steps(double* out, int cap) {
    int n = (int)g_steps.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = g_steps[i];
    g_steps.clear();
    return n;
}

void* hs_model_create(const sfx_model_desc* desc, char* err, int errlen) {
    SimHandle* h = new SimHandle();
    h->use_double = desc->use_double;
    std::string e = desc->use_double ? prepare_model(*desc, h->d.h) : prepare_model(*desc, h->f.h);
    if (!e.empty()) {
        std::strncpy(err, e.c_str(), errlen - 1);
        delete h;
        return nullptr;
    }
    return h;
}
void hs_model_destroy(void* p) { delete (SimHandle*)p; }

void hs_set_gmm(void* p, int M, int D, const void* means, const void* prec, const void* logw) {
    SimHandle* h = (SimHandle*)p;
    if (h->use_double) set_gmm<double>(h->d, M, D, (const double*)means, (const double*)prec, (const double*)logw);
    else set_gmm<float>(h->f, M, D, (const float*)means, (const float*)prec, (const float*)logw);
}

int hs_layout(void* p, int use_vposer, SfxLayout* out) {
    SimHandle* h = (SimHandle*)p;
    if (h->use_double) *out = make_layout(h->d.h.NB, h->d.h.NE, h->d.h.NH, use_vposer);
    else *out = make_layout(h->f.h.NB, h->f.h.NE, h->f.h.NH, use_vposer);
    return 0;
}

// one frame; all arrays in the model dtype
int hs_run(void* p, const SfxStage* st, int use_vposer, int do_fit, void* params, const void* gt,
           const void* conf, const void* jw, const unsigned char* lowconf,
           const unsigned char* init_mask, const void* cam, const void* reg_pose, void* loss_out,
           void* grad_out, void* joints_out, int* n_evals, int* flags) {
    SimHandle* h = (SimHandle*)p;
    if (h->use_double)
        run<double>(h->d, st, use_vposer, do_fit, (double*)params, (const double*)gt,
                    (const double*)conf, (const double*)jw, lowconf, init_mask, (const double*)cam,
                    (const double*)reg_pose, (double*)loss_out, (double*)grad_out,
                    (double*)joints_out, n_evals, flags);
    else
        run<float>(h->f, st, use_vposer, do_fit, (float*)params, (const float*)gt,
                   (const float*)conf, (const float*)jw, lowconf, init_mask, (const float*)cam,
                   (const float*)reg_pose, (float*)loss_out, (float*)grad_out, (float*)joints_out,
                   n_evals, flags);
    return 0;
}
}
