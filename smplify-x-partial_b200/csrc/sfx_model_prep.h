// Host-side re-layout of the raw SMPL-X arrays (sfx_model_desc) into the tables the kernels
// read.  Pure C++ (no CUDA) so the library and the host-simulation tests share it.
#pragma once
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/sfx.h"
#include "sfx_core.cuh"

namespace sfx {

template <typename T>
struct HostModel {
    int V = 0, F = 0, NS = 0, NB = 0, NE = 0, NH = 0, K = 0, NJOUT = 0, use_contour = 0;
    int wide_model = 0;      // the model arrays were given in float64 and T is double: no float32 caches
    std::vector<T> PK, vt, J0, JS, Wd, hand_l, hand_r, pose_mean, lmk_bary, dyn_bary;
    std::vector<int> sv_vid, dyn_vid, joint_map, inv_ptr, inv_idx, faces, sk_ptr;
    std::vector<unsigned char> sk_j;
    std::vector<T> sk_w;
    std::vector<DynRowPack<T>> dyn_pack;     // support tables of the dynamic-contour slots per yaw row
    int parents[SFX_NJ], order[SFX_NJ], level_off[16], nlev = 0;
    int child_off[SFX_NJ + 1], child_idx[SFX_NJ], neck[8], n_neck = 0;

    // fills everything except the array pointers (those point at host or device copies)
    void fill_scalars(ModelView<T>& m) const {
        m.V = V; m.NS = NS; m.NB = NB; m.NE = NE; m.NH = NH; m.K = K; m.NJOUT = NJOUT;
        m.use_contour = use_contour; m.n_neck = n_neck; m.nlev = nlev;
        m.wide_model = wide_model;
        m.vp_ready = 0; m.vp_w1 = m.vp_b1 = m.vp_w2 = m.vp_b2 = m.vp_w3 = m.vp_b3 = nullptr;
        m.gmm_M = 0; m.gmm_D = 0; m.gmm_means = nullptr; m.gmm_prec = nullptr; m.gmm_logw = nullptr;
        m.coll_ready = 0; m.F = F; m.n_parts = 0; m.faces = nullptr; m.part_ptr = nullptr;
        m.part_faces = nullptr; m.n_clusters = 0; m.cl_ptr = nullptr; m.part_cl_ptr = nullptr;
        m.face_part = nullptr; m.part_allow = nullptr; m.vf_ptr = nullptr;
        m.vf_idx = nullptr;
        m.dyn_pack = nullptr;
        for (int i = 0; i < SFX_NJ; ++i) { m.parents[i] = parents[i]; m.order[i] = order[i]; }
        for (int i = 0; i < 16; ++i) m.level_off[i] = level_off[i];
        for (int i = 0; i <= SFX_NJ; ++i) m.child_off[i] = child_off[i];
        for (int i = 0; i < SFX_NJ; ++i) m.child_idx[i] = child_idx[i];
        for (int i = 0; i < 8; ++i) m.neck[i] = neck[i];
    }
    ModelView<T> host_view() const {
        ModelView<T> m;
        fill_scalars(m);
        m.PK = PK.data(); m.vt = vt.data(); m.J0 = J0.data(); m.JS = JS.data(); m.Wd = Wd.data();
        m.hand_l = hand_l.data(); m.hand_r = hand_r.data(); m.pose_mean = pose_mean.data();
        m.sv_vid = sv_vid.data(); m.lmk_bary = lmk_bary.data(); m.dyn_vid = dyn_vid.data();
        m.dyn_bary = dyn_bary.data(); m.joint_map = joint_map.data();
        m.inv_ptr = inv_ptr.data(); m.inv_idx = inv_idx.data();
        m.sk_ptr = sk_ptr.data(); m.sk_j = sk_j.data(); m.sk_w = sk_w.data();
        m.dyn_pack = dyn_pack.empty() ? nullptr : dyn_pack.data();
        return m;
    }
};

template <typename T>
std::string prepare_model(const sfx_model_desc& d, HostModel<T>& h) {
    const int V = d.num_verts;
    if (V <= 0 || !d.v_template || !d.shapedirs || !d.posedirs || !d.J_regressor ||
        !d.lbs_weights || !d.parents || !d.faces || !d.joint_map || !d.extra_vertex_ids ||
        !d.lmk_faces_idx || !d.lmk_bary_coords)
        return "sfx_model_desc: missing array";
    if (d.num_betas + d.num_expr > SFX_NSHAPE_MAX) return "num_betas + num_expr exceeds 26";
    if (d.n_hand < 1 || d.n_hand > 45) return "n_hand must be in 1..45";
    if (d.num_keypoints < 1 || d.num_keypoints > SFX_KMAX) return "num_keypoints out of range";
    if (d.use_face_contour && (!d.dyn_lmk_faces_idx || !d.dyn_lmk_bary_coords))
        return "use_face_contour needs the dynamic landmark tables";
    // floating-point arrays of the description: float32 or float64 (arrays_float64)
    const bool f64 = d.arrays_float64 != 0;
    auto rd = [f64](const void* p, size_t i) -> double {
        return f64 ? static_cast<const double*>(p)[i] : (double)static_cast<const float*>(p)[i];
    };
    h.wide_model = (f64 && sizeof(T) == 8) ? 1 : 0;
    h.V = V; h.F = d.num_faces; h.NB = d.num_betas; h.NE = d.num_expr; h.NS = h.NB + h.NE;
    h.NH = d.n_hand; h.K = d.num_keypoints; h.use_contour = d.use_face_contour ? 1 : 0;
    h.NJOUT = SFX_NJ + SFX_NEXTRA + SFX_NLMK + (h.use_contour ? SFX_NDYN : 0);
    const int NS = h.NS;
    // --- kinematic tree ---
    int depth[SFX_NJ];
    for (int j = 0; j < SFX_NJ; ++j) {
        h.parents[j] = d.parents[j];
        if (j == 0 ? h.parents[j] >= 0 : (h.parents[j] < 0 || h.parents[j] >= j))
            return "parents must be a topologically ordered tree rooted at joint 0";
        depth[j] = j == 0 ? 0 : depth[h.parents[j]] + 1;
    }
    int maxd = *std::max_element(depth, depth + SFX_NJ);
    if (maxd + 1 > 15) return "kinematic tree too deep";
    h.nlev = maxd + 1;
    int pos = 0;
    for (int lv = 0; lv <= maxd; ++lv) {
        h.level_off[lv] = pos;
        for (int j = 0; j < SFX_NJ; ++j)
            if (depth[j] == lv) h.order[pos++] = j;
    }
    for (int lv = maxd + 1; lv < 16; ++lv) h.level_off[lv] = pos;
    pos = 0;
    for (int j = 0; j < SFX_NJ; ++j) {
        h.child_off[j] = pos;
        for (int c = 0; c < SFX_NJ; ++c)
            if (h.parents[c] == j) h.child_idx[pos++] = c;
    }
    h.child_off[SFX_NJ] = pos;
    for (; pos < SFX_NJ; ++pos) h.child_idx[pos] = 0;
    h.n_neck = 0;
    for (int cur = 12; cur != -1 && h.n_neck < 8; cur = h.parents[cur]) h.neck[h.n_neck++] = cur;
    for (int i = h.n_neck; i < 8; ++i) h.neck[i] = 0;
    // --- blend matrix PK [3V][512]: pose dirs | shape dirs | 0 ---
    h.PK.assign((size_t)3 * V * SFX_KPAD, (T)0);
    h.vt.resize((size_t)3 * V);
    for (long r = 0; r < 3L * V; ++r) {
        T* row = h.PK.data() + r * SFX_KPAD;
        const size_t pd = (size_t)r * SFX_NPF, sd = (size_t)r * d.shape_stride;
        for (int k = 0; k < SFX_NPF; ++k) row[k] = (T)rd(d.posedirs, pd + k);
        for (int s = 0; s < h.NB; ++s) row[SFX_NPF + s] = (T)rd(d.shapedirs, sd + s);
        for (int s = 0; s < h.NE; ++s) row[SFX_NPF + h.NB + s] = (T)rd(d.shapedirs, sd + d.expr_offset + s);
        h.vt[r] = (T)rd(d.v_template, r);
    }
    // --- rest joints: J0 = Jreg . v_template, JS = Jreg . shapedirs (accumulated in double) ---
    h.J0.assign(SFX_NJ * 3, (T)0);
    h.JS.assign(SFX_NJ * 3 * 32, (T)0);
    {
        std::vector<double> acc(SFX_NJ * 3 * 33);
        std::fill(acc.begin(), acc.end(), 0.0);
        for (int j = 0; j < SFX_NJ; ++j) {
            for (int v = 0; v < V; ++v) {
                double w = rd(d.J_regressor, (size_t)j * V + v);
                if (w == 0.0) continue;
                for (int c = 0; c < 3; ++c) {
                    long r = 3L * v + c;
                    double* a = acc.data() + (j * 3 + c) * 33;
                    a[32] += w * rd(d.v_template, r);
                    const size_t sd = (size_t)r * d.shape_stride;
                    for (int s = 0; s < h.NB; ++s) a[s] += w * rd(d.shapedirs, sd + s);
                    for (int s = 0; s < h.NE; ++s) a[h.NB + s] += w * rd(d.shapedirs, sd + d.expr_offset + s);
                }
            }
        }
        for (int i = 0; i < SFX_NJ * 3; ++i) {
            h.J0[i] = (T)acc[i * 33 + 32];
            for (int s = 0; s < NS; ++s) h.JS[i * 32 + s] = (T)acc[i * 33 + s];
        }
    }
    // --- dense skinning weights, padded rows ---
    h.Wd.assign((size_t)V * SFX_WROW, (T)0);
    for (long v = 0; v < V; ++v)
        for (int j = 0; j < SFX_NJ; ++j) h.Wd[v * SFX_WROW + j] = (T)rd(d.lbs_weights, v * SFX_NJ + j);
    h.sk_ptr.assign(V + 1, 0);
    h.sk_j.clear();
    h.sk_w.clear();
    for (long v = 0; v < V; ++v) {
        for (int j = 0; j < SFX_NJ; ++j)
            if (rd(d.lbs_weights, v * SFX_NJ + j) != 0.0) {
                h.sk_j.push_back((unsigned char)j);
                h.sk_w.push_back((T)rd(d.lbs_weights, v * SFX_NJ + j));
            }
        h.sk_ptr[v + 1] = (int)h.sk_j.size();
    }
    // --- hands ---
    h.hand_l.resize((size_t)h.NH * 45);
    h.hand_r.resize((size_t)h.NH * 45);
    for (int i = 0; i < h.NH * 45; ++i) {
        h.hand_l[i] = (T)rd(d.hand_components_l, i);
        h.hand_r[i] = (T)rd(d.hand_components_r, i);
    }
    h.pose_mean.assign(SFX_NPOSE, (T)0);
    for (int i = 0; i < 45; ++i) {
        h.pose_mean[75 + i] = d.hand_mean_l ? (T)rd(d.hand_mean_l, i) : (T)0;
        h.pose_mean[120 + i] = d.hand_mean_r ? (T)rd(d.hand_mean_r, i) : (T)0;
    }
    // --- support vertices ---
    h.faces.resize((size_t)3 * h.F);
    for (long i = 0; i < 3L * h.F; ++i) {
        h.faces[i] = d.faces[i];
        if (h.faces[i] < 0 || h.faces[i] >= V) return "face index out of range";
    }
    h.sv_vid.resize(SFX_NSTATIC);
    h.lmk_bary.resize(SFX_NLMK * 3);
    for (int i = 0; i < SFX_NEXTRA; ++i) {
        h.sv_vid[i] = d.extra_vertex_ids[i];
        if (h.sv_vid[i] < 0 || h.sv_vid[i] >= V) return "extra vertex id out of range";
    }
    for (int l = 0; l < SFX_NLMK; ++l) {
        int f = d.lmk_faces_idx[l];
        if (f < 0 || f >= h.F) return "lmk_faces_idx out of range";
        for (int k = 0; k < 3; ++k) {
            h.sv_vid[SFX_NEXTRA + 3 * l + k] = h.faces[3L * f + k];
            h.lmk_bary[3 * l + k] = (T)rd(d.lmk_bary_coords, 3 * l + k);
        }
    }
    h.dyn_vid.assign(SFX_NDYNROWS * 51, 0);
    h.dyn_bary.assign(SFX_NDYNROWS * 51, (T)0);
    if (h.use_contour) {
        for (int y = 0; y < SFX_NDYNROWS; ++y)
            for (int i = 0; i < SFX_NDYN; ++i) {
                int f = d.dyn_lmk_faces_idx[y * SFX_NDYN + i];
                if (f < 0 || f >= h.F) return "dynamic_lmk_faces_idx out of range";
                for (int k = 0; k < 3; ++k) {
                    h.dyn_vid[y * 51 + 3 * i + k] = h.faces[3L * f + k];
                    h.dyn_bary[y * 51 + 3 * i + k] =
                        (T)rd(d.dyn_lmk_bary_coords, (size_t)(y * SFX_NDYN + i) * 3 + k);
                }
            }
    }
    // --- support tables of the dynamic-contour slots, one per yaw row: what the evaluation's
    //     support_slots() + support_by_joint() would build (sfx_core.cuh), done once here ---
    {
        const int rows = h.use_contour ? SFX_NDYNROWS : 1;
        h.dyn_pack.assign(rows, DynRowPack<T>());
        for (int y = 0; y < rows; ++y) {
            DynRowPack<T>& P = h.dyn_pack[y];
            memset(&P, 0, sizeof(P));
            for (int i = 0; i < SFX_NDYNSLOT; ++i) {
                const int vid = h.use_contour ? h.dyn_vid[y * 51 + i] : h.sv_vid[0];
                P.vid[i] = vid;
                P.bary[i] = h.use_contour ? h.dyn_bary[y * 51 + i] : (T)0;
                for (int k = 0; k < 3; ++k) P.vt_s[3 * i + k] = h.vt[3L * vid + k];
                const int e0 = h.sk_ptr[vid], n = h.sk_ptr[vid + 1] - e0;
                for (int e = 0; e < n && e < SFX_NW; ++e) {
                    P.wj[i * SFX_NW + e] = h.sk_j[e0 + e];
                    P.ww[i * SFX_NW + e] = (float)h.sk_w[e0 + e];
                }
                P.wn[i] = (unsigned char)(n <= SFX_NW ? n : SFX_NW);
                if (n > SFX_NW) P.overflow = 1;
            }
            int pos = SFX_NSTATIC * SFX_NW;
            for (int j = 0; j < SFX_NJ; ++j) {
                P.jtd_ptr[j] = pos;
                for (int i = 0; i < SFX_NDYNSLOT; ++i)
                    for (int e = 0; e < P.wn[i]; ++e)
                        if (P.wj[i * SFX_NW + e] == j) {
                            P.jt_slot[pos - SFX_NSTATIC * SFX_NW] = (unsigned char)(SFX_NSTATIC + i);
                            P.jt_w[pos - SFX_NSTATIC * SFX_NW] = P.ww[i * SFX_NW + e];
                            ++pos;
                        }
            }
            P.jtd_ptr[SFX_NJ] = pos;
        }
    }
    // --- joint mapper and its inverse ---
    h.joint_map.resize(h.K);
    std::vector<int> cnt(h.NJOUT + 1, 0);
    for (int k = 0; k < h.K; ++k) {
        int j = d.joint_map[k];
        if (j < 0 || j >= h.NJOUT) return "joint_map entry out of range";
        h.joint_map[k] = j;
        cnt[j + 1]++;
    }
    h.inv_ptr.assign(h.NJOUT + 1, 0);
    for (int j = 0; j < h.NJOUT; ++j) h.inv_ptr[j + 1] = h.inv_ptr[j] + cnt[j + 1];
    h.inv_idx.resize(h.K);
    std::vector<int> fillp(h.inv_ptr.begin(), h.inv_ptr.end() - 1);
    for (int k = 0; k < h.K; ++k) h.inv_idx[fillp[h.joint_map[k]]++] = k;
    return "";
}

// Tables of the interpenetration term (sfx_collide.cuh) from the face segmentation the reference
// reads at fit_single_frame.py:317-328 (smplx_parts_segm.pkl: 'segm' = body part of every face,
// 'parents' = kinematic parent of that part) and its ign_part_pairs option.
struct HostCollision {
    int n_parts = 0, n_clusters = 0;
    std::vector<int> part_ptr, part_faces, vf_ptr, vf_idx, cl_ptr, part_cl_ptr;
    std::vector<unsigned char> face_part;
    std::vector<unsigned long long> part_allow;
};

#define SFX_CLUSTER_FACES 64      // faces per cluster of the broad phase
#define SFX_NCLUSTER_MAX 512

template <typename TV>
inline std::string prepare_collision(int V, int F, const int* faces, const TV* vt, const int32_t* segm,
                                     const int32_t* parents, const int32_t* ign_pairs, int n_ign,
                                     HostCollision& c, bool unfiltered = false) {
    // unfiltered: no FilterFaces (fit_single_frame.py:317-328 with an empty part_segm_fn) -- `segm`
    // only groups the faces for the broad phase and every pair of groups is admissible, a group
    // with itself included; faces that share a vertex are sorted out by the narrow phase
    std::vector<int32_t> no_parents;
    if (unfiltered && !parents && segm && F >= 1) {
        no_parents.assign(F, -1);
        parents = no_parents.data();
    }
    if (!faces || !segm || !parents || F < 1) return "collision tables: missing array";
    if (F > 65535) return "collision tables: more than 65535 faces";
    int np = 0;
    for (int f = 0; f < F; ++f) {
        if (segm[f] < 0 || segm[f] >= SFX_NPART_MAX) return "faces_segm entry out of range (0..63)";
        if (parents[f] >= SFX_NPART_MAX) return "faces_parents entry out of range";
        np = std::max(np, segm[f] + 1);
    }
    c.n_parts = np;
    std::vector<int> parent_of(np, -1);
    for (int f = 0; f < F; ++f) parent_of[segm[f]] = parents[f];
    // FilterFaces: same part, parent / child parts and ignored pairs never collide
    c.part_allow.assign(SFX_NPART_MAX, 0ull);
    for (int p = 0; p < np; ++p)
        for (int q = 0; q < np; ++q)
            if (unfiltered || (p != q && parent_of[p] != q && parent_of[q] != p)) c.part_allow[p] |= 1ull << q;
    for (int i = 0; i < n_ign; ++i) {
        const int a = ign_pairs[2 * i], b = ign_pairs[2 * i + 1];
        if (a < 0 || b < 0) return "ign_part_pairs entry out of range";
        if (a >= np || b >= np) continue;
        c.part_allow[a] &= ~(1ull << b);
        c.part_allow[b] &= ~(1ull << a);
    }
    c.face_part.resize(F);
    c.part_ptr.assign(np + 1, 0);
    for (int f = 0; f < F; ++f) {
        c.face_part[f] = (unsigned char)segm[f];
        c.part_ptr[segm[f] + 1]++;
    }
    for (int p = 0; p < np; ++p) c.part_ptr[p + 1] += c.part_ptr[p];
    c.part_faces.resize(F);
    {
        std::vector<int> pos(c.part_ptr.begin(), c.part_ptr.end() - 1);
        for (int f = 0; f < F; ++f) c.part_faces[pos[segm[f]]++] = f;
    }
    // Clusters: inside a part the faces are ordered along the longest side of the part's
    // template box and cut into runs of SFX_CLUSTER_FACES -- a static two-level hierarchy whose
    // boxes are refitted at every evaluation (the broad phase of sfx_collide.cuh).
    c.cl_ptr.clear();
    c.part_cl_ptr.assign(np + 1, 0);
    for (int p = 0; p < np; ++p) {
        const int a = c.part_ptr[p], b = c.part_ptr[p + 1];
        if (b > a) {
            double lo[3] = {1e30, 1e30, 1e30}, hi[3] = {-1e30, -1e30, -1e30};
            std::vector<std::pair<double, int>> key(b - a);
            std::vector<double> cen((size_t)(b - a) * 3);
            for (int k = a; k < b; ++k) {
                const int f = c.part_faces[k];
                for (int d = 0; d < 3; ++d) {
                    double s = 0;
                    for (int e = 0; e < 3; ++e) s += (double)vt[3L * faces[3 * f + e] + d];
                    cen[(size_t)(k - a) * 3 + d] = s / 3;
                    lo[d] = std::min(lo[d], s / 3);
                    hi[d] = std::max(hi[d], s / 3);
                }
            }
            int ax = 0;
            for (int d = 1; d < 3; ++d)
                if (hi[d] - lo[d] > hi[ax] - lo[ax]) ax = d;
            for (int k = a; k < b; ++k) key[k - a] = {cen[(size_t)(k - a) * 3 + ax], c.part_faces[k]};
            std::sort(key.begin(), key.end());
            for (int k = a; k < b; ++k) c.part_faces[k] = key[k - a].second;
            for (int k = a; k < b; k += SFX_CLUSTER_FACES) c.cl_ptr.push_back(k);
        }
        c.part_cl_ptr[p + 1] = (int)c.cl_ptr.size();
    }
    c.n_clusters = (int)c.cl_ptr.size();
    c.cl_ptr.push_back(F);
    if (c.n_clusters > SFX_NCLUSTER_MAX) return "collision tables: too many face clusters";
    // cluster boundaries must not cross parts: the sentinel of a part's last cluster is the
    // next part's first face, which is what cl_ptr holds by construction
    c.vf_ptr.assign(V + 1, 0);
    for (long i = 0; i < 3L * F; ++i) c.vf_ptr[faces[i] + 1]++;
    for (int v = 0; v < V; ++v) c.vf_ptr[v + 1] += c.vf_ptr[v];
    c.vf_idx.resize((size_t)3 * F);
    {
        std::vector<int> pos(c.vf_ptr.begin(), c.vf_ptr.end() - 1);
        for (int f = 0; f < F; ++f)
            for (int k = 0; k < 3; ++k) c.vf_idx[pos[faces[3 * f + k]]++] = f * 4 + k;
    }
    return "";
}

inline SfxLayout make_layout(int n_betas, int n_expr, int n_hand, int use_vposer) {
    SfxLayout L;
    L.n_betas = n_betas; L.n_expr = n_expr; L.n_hand = n_hand;
    L.n_pose = use_vposer ? SFX_NLATENT : 63;
    int o = 0;
    L.off_betas = o; o += n_betas;
    L.off_go = o; o += 3;
    L.off_lh = o; o += n_hand;
    L.off_rh = o; o += n_hand;
    L.off_jaw = o; o += 3;
    L.off_leye = o; o += 3;
    L.off_reye = o; o += 3;
    L.off_expr = o; o += n_expr;
    L.off_pose = o; o += L.n_pose;
    L.off_camt = o; o += 3;
    L.np = o;
    return L;
}

}  // namespace sfx
