// Interpenetration term of SMPLifyLoss (reference fitting.py:437-455) for one frame, inside the
// per-frame evaluation: self-collision search over the F = 20 908 triangles of the posed mesh,
// body-part filter, conic distance-field penalty and its analytic adjoint.
//
// The reference delegates all three steps to the un-vendored third-party package
// mesh_intersection (BVH search tree, FilterFaces, DistanceFieldPenetrationLoss: call sites
// fit_single_frame.py:301-328, fitting.py:441-455).  What is restated here is the published
// algorithm (see oracle/isect_port.py for the formulas and the citation); parity is unpinned.
//
// B200 design.  One block owns one frame, so the search is built for 512 threads and the
// shared memory of one SM instead of a device-wide LBVH:
//   * the part filter comes FIRST: only faces of parts (p, q) that FilterFaces would keep can
//     form a pair, so the broad phase works on a static two-level hierarchy (faces grouped by
//     part and, inside a part, cut into clusters of 64 along the part's long axis) whose boxes
//     are refitted per evaluation; a face is a "candidate" iff its box reaches a cluster box of
//     an admissible partner part;
//   * candidates are compacted in a fixed order and sorted by (extent class, box minimum along
//     the body's longest axis) in the shared-memory area that holds the blend-row ring during the
//     streaming passes, each with a packed 8-byte entry (part + box on a 256-level grid); a warp
//     per candidate walks the windows of the four class runs that can overlap it and lists the
//     box-overlapping, admissible partners;
//   * the narrow phase GATHERS: the warp owning a candidate runs the separating-axis test against
//     its listed partners and accumulates the penalty of the candidate's own cone plus the
//     gradient w.r.t. its own three corners -- no atomics on values, fixed summation order, so
//     the result is bit-reproducible run to run; the per-vertex sums walk the static
//     vertex->face table in order for the same reason.
// The same source compiles for the single-threaded host simulation (tests only).
#pragma once

namespace sfx {

// per-block workspace of the term
template <typename T>
struct CollWS {
    T* vp_g;               // [3V]  blended (unposed) vertices of the current evaluation
    T* vert_g;             // [3V]  posed vertices; after the search: dL/dv_posed of touched vertex t at [3 t ..]
    T* dvert_g;            // [3V]  dL/dvertex of touched vertices
    T* dtri_g;             // [F][9] dL/d(own corners) of faces that collided
    unsigned short* tv_g;  // [V]   touched vertices of the current evaluation, ascending
    T* fbox;               // [F][6] axis-aligned box of every face
    T* ftri;               // [F][9] its three corners, contiguous (one 36-byte read per partner)
    unsigned char* sort_g; // [SFX_COLL_ENTRY * SFX_COLL_SORT_G] sweep arrays when they outgrow the shared area
    unsigned short* hits_g;     // [warps][32 * hits_cap] per-warp lists: (count, partner faces...) per candidate
    int hits_cap;
    unsigned char* work;   // shared-memory work area (the idle blend ring on the device)
    int work_bytes;
};
#define SFX_COLL_SORT_G 32768      // capacity of the global sweep arrays (power of two >= F)
#define SFX_COLL_CLASSES 4         // extent classes of the sweep (quantised extent <= 4, 16, 64, any)
#define SFX_COLL_ENTRY 12          // bytes per candidate: packed part + quantised box 8, key 2, face 2
#define SFX_COLL_HITS 2048         // potential-hit list: 16-bit entries per lane (a warp's region is 32 x this)

template <typename T>
SFX_FN CollWS<T> coll_block_ws(int V, int F, T* vals, unsigned short* idx, unsigned char* work,
                               int work_bytes) {
    CollWS<T> W;
    W.vp_g = vals;
    W.vert_g = vals + 3L * V;
    W.dvert_g = vals + 6L * V;
    W.dtri_g = vals + 9L * V;
    W.fbox = vals + 9L * V + 9L * F;
    W.ftri = vals + 9L * V + 15L * F;
    W.tv_g = idx;
    W.sort_g = reinterpret_cast<unsigned char*>(idx + (V + 7) / 8 * 8);
    W.hits_g = idx + (V + 7) / 8 * 8 + (long)SFX_COLL_ENTRY * SFX_COLL_SORT_G / 2;
    W.hits_cap = SFX_COLL_HITS;
    W.work = work;
    W.work_bytes = work_bytes;
    return W;
}
inline long coll_vals_per_block(int V, int F) { return (9L * V + 24L * F + 3) / 4 * 4; }
inline long coll_idx_per_block(int V, int F) { return ((long)V + 7) / 8 * 8 + (long)SFX_COLL_ENTRY * SFX_COLL_SORT_G / 2 + 512L * SFX_COLL_HITS; }

template <typename T>
SFX_FN void v3sub(const T* a, const T* b, T* c) { c[0] = a[0] - b[0]; c[1] = a[1] - b[1]; c[2] = a[2] - b[2]; }
template <typename T>
SFX_FN void v3cross(const T* a, const T* b, T* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
template <typename T>
SFX_FN T v3dot(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// ---- separating-axis test, 17 axes (2 normals, 9 edge x edge, 6 in-plane edge normals) -------
template <typename T>
SFX_FN bool axis_separates(const T* ax, const T* t1, const T* t2) {
    T a0 = v3dot(ax, t1), a1 = v3dot(ax, t1 + 3), a2 = v3dot(ax, t1 + 6);
    T b0 = v3dot(ax, t2), b1 = v3dot(ax, t2 + 3), b2 = v3dot(ax, t2 + 6);
    T amin = a0 < a1 ? a0 : a1; amin = amin < a2 ? amin : a2;
    T amax = a0 > a1 ? a0 : a1; amax = amax > a2 ? amax : a2;
    T bmin = b0 < b1 ? b0 : b1; bmin = bmin < b2 ? bmin : b2;
    T bmax = b0 > b1 ? b0 : b1; bmax = bmax > b2 ? bmax : b2;
    return amax < bmin || bmax < amin;
}

template <typename T>
SFX_FN_NOINLINE bool triangles_intersect(const T* t1, const T* t2) {
    T e1[9], e2[9], n1[3], n2[3], ax[3];
    v3sub(t1 + 3, t1, e1); v3sub(t1 + 6, t1, e1 + 3); v3sub(t1 + 6, t1 + 3, e1 + 6);
    v3sub(t2 + 3, t2, e2); v3sub(t2 + 6, t2, e2 + 3); v3sub(t2 + 6, t2 + 3, e2 + 6);
    v3cross(e1, e1 + 3, n1);
    v3cross(e2, e2 + 3, n2);
    if (axis_separates(n1, t1, t2)) return false;
    if (axis_separates(n2, t1, t2)) return false;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            v3cross(e1 + 3 * i, e2 + 3 * j, ax);
            if (axis_separates(ax, t1, t2)) return false;
        }
    for (int i = 0; i < 3; ++i) {
        v3cross(n1, e1 + 3 * i, ax);
        if (axis_separates(ax, t1, t2)) return false;
    }
    for (int j = 0; j < 3; ++j) {
        v3cross(n2, e2 + 3 * j, ax);
        if (axis_separates(ax, t1, t2)) return false;
    }
    return true;
}

// ---- cone of a triangle: circumcentre o, circumradius r, unit normal n ------------------------
template <typename T>
struct Cone {
    T a[3], b[3], cr[3], w[3], m[3], o[3], n[3];
    T cc, aa, bb, e, sq, rc, r;
};

template <typename T>
SFX_FN void cone_make(const T* tri, Cone<T>& C) {
    v3sub(tri, tri + 6, C.a);
    v3sub(tri + 3, tri + 6, C.b);
    v3cross(C.a, C.b, C.cr);
    C.cc = v3dot(C.cr, C.cr);
    C.aa = v3dot(C.a, C.a);
    C.bb = v3dot(C.b, C.b);
    T ab[3];
    v3sub(C.a, C.b, ab);
    C.e = v3dot(ab, ab);
    C.sq = sfx_sqrt(C.e * C.aa * C.bb);
    C.rc = sfx_sqrt(C.cc);
    C.r = C.sq / ((T)2 * C.rc);
    for (int k = 0; k < 3; ++k) C.w[k] = C.aa * C.b[k] - C.bb * C.a[k];
    v3cross(C.w, C.cr, C.m);
    for (int k = 0; k < 3; ++k) {
        C.o[k] = tri[6 + k] + C.m[k] / ((T)2 * C.cc);
        C.n[k] = C.cr[k] / C.rc;
    }
}

// (gn, go, gr) = dL/d(normal, centre, radius)  ->  g[9] += dL/d(corners)
template <typename T>
SFX_FN void cone_backward(const Cone<T>& C, const T* gn, const T* go, T gr, T* g) {
    T ga[3] = {0, 0, 0}, gb[3] = {0, 0, 0}, gcr[3], gw[3], gm[3], tmp[3];
    T gcc = -gr * C.r / ((T)2 * C.cc);
    const T gq = gr / ((T)2 * C.rc) / ((T)2 * C.sq);
    const T s = (T)1 / ((T)2 * C.cc);
    for (int k = 0; k < 3; ++k) gm[k] = go[k] * s;
    gcc += -v3dot(go, C.m) / ((T)2 * C.cc * C.cc);
    for (int k = 0; k < 3; ++k) gcr[k] = gn[k] / C.rc;
    gcc += -v3dot(gn, C.cr) / ((T)2 * C.cc * C.rc);
    v3cross(C.cr, gm, gw);                       // m = w x cr
    v3cross(gm, C.w, tmp);
    for (int k = 0; k < 3; ++k) gcr[k] += tmp[k];
    T gaa = v3dot(gw, C.b), gbb = -v3dot(gw, C.a);
    for (int k = 0; k < 3; ++k) { gb[k] += C.aa * gw[k]; ga[k] += -C.bb * gw[k]; }
    const T ge = gq * C.aa * C.bb;
    gaa += gq * C.e * C.bb;
    gbb += gq * C.e * C.aa;
    for (int k = 0; k < 3; ++k) {
        const T gab = (T)2 * ge * (C.a[k] - C.b[k]);
        ga[k] += gab;
        gb[k] -= gab;
    }
    for (int k = 0; k < 3; ++k) gcr[k] += (T)2 * gcc * C.cr[k];
    for (int k = 0; k < 3; ++k) { ga[k] += (T)2 * gaa * C.a[k]; gb[k] += (T)2 * gbb * C.b[k]; }
    v3cross(C.b, gcr, tmp);                      // cr = a x b
    for (int k = 0; k < 3; ++k) ga[k] += tmp[k];
    v3cross(gcr, C.a, tmp);
    for (int k = 0; k < 3; ++k) gb[k] += tmp[k];
    for (int k = 0; k < 3; ++k) {
        g[k] += ga[k];
        g[3 + k] += gb[k];
        g[6 + k] += go[k] - ga[k] - gb[k];
    }
}

// psi(v) = (1 - Phi) Upsilon for one point; with the partial results the adjoint needs
template <typename T>
struct Field {
    T d[3], p[3], x, rad, den, phi, ups, dups, psi;
    bool live;
};

template <typename T>
SFX_FN void cone_field(const Cone<T>& C, const T* v, T sigma, Field<T>& F) {
    v3sub(v, C.o, F.d);
    F.x = v3dot(F.d, C.n);
    for (int k = 0; k < 3; ++k) F.p[k] = F.d[k] - F.x * C.n[k];
    F.rad = sfx_sqrt(v3dot(F.p, F.p));
    F.den = C.r * ((T)1 - F.x / sigma);
    F.phi = F.rad / F.den;
    if (F.x <= -sigma) {
        F.ups = -F.x + (T)1 - sigma;
        F.dups = (T)-1;
    } else if (F.x < sigma) {
        const T q2 = -((T)1 - (T)2 * sigma) / ((T)4 * sigma * sigma), q1 = -(T)1 / ((T)2 * sigma);
        F.ups = q2 * F.x * F.x + q1 * F.x + ((T)3 - (T)2 * sigma) / (T)4;
        F.dups = (T)2 * q2 * F.x + q1;
    } else {
        F.ups = 0;
        F.dups = 0;
    }
    F.live = ((F.phi) < (T)1) && ((F.den) > (T)0);
    F.psi = F.live ? ((T)1 - F.phi) * F.ups : (T)0;
}

// gpsi = dL/dpsi -> gv[3] += dL/dv and, when cone gradients are wanted, gn / go / gr
template <typename T>
SFX_FN void cone_field_backward(const Cone<T>& C, const Field<T>& F, T sigma, T gpsi, T* gv,
                                T* gn, T* go, T* gr) {
    if (!F.live) return;
    const T gphi = -gpsi * F.ups;
    T gx = gpsi * ((T)1 - F.phi) * F.dups;
    const T grad = gphi / F.den;
    const T gden = -gphi * F.rad / (F.den * F.den);
    if (gr) *gr += gden * ((T)1 - F.x / sigma);
    gx += gden * (-C.r / sigma);
    T gp[3] = {0, 0, 0};
    if (F.rad > (T)0)
        for (int k = 0; k < 3; ++k) gp[k] = grad * F.p[k] / F.rad;
    gx += -v3dot(gp, C.n);
    T gd[3];
    for (int k = 0; k < 3; ++k) gd[k] = gp[k] + gx * C.n[k];
    if (gn)
        for (int k = 0; k < 3; ++k) gn[k] += -F.x * gp[k] + gx * F.d[k];
    if (gv)
        for (int k = 0; k < 3; ++k) gv[k] += gd[k];
    if (go)
        for (int k = 0; k < 3; ++k) go[k] -= gd[k];
}

// Everything the owner of triangle i takes from the colliding pair (i, j):
//   loss += sum_{v in j} Psi_i(v)^2                         (the cone of i; Psi = psi^2)
//   gi   += d/d(corners of i) [ sum_{v in j} Psi_i(v)^2 + sum_{v in i} Psi_j(v)^2 ]
template <typename T>
SFX_FN_NOINLINE void pair_terms(const T* ti, const T* tj, T sigma, T* loss, T* gi) {
    Cone<T> C;
    Field<T> F;
    cone_make(ti, C);
    T gn[3] = {0, 0, 0}, go[3] = {0, 0, 0}, gr = 0, acc = 0;
    for (int k = 0; k < 3; ++k) {
        cone_field(C, tj + 3 * k, sigma, F);
        const T p2 = F.psi * F.psi;
        acc += p2 * p2;
        cone_field_backward<T>(C, F, sigma, (T)4 * p2 * F.psi, nullptr, gn, go, &gr);
    }
    cone_backward(C, gn, go, gr, gi);
    *loss += acc;
    cone_make(tj, C);
    for (int k = 0; k < 3; ++k) {
        cone_field(C, ti + 3 * k, sigma, F);
        const T p2 = F.psi * F.psi;
        cone_field_backward<T>(C, F, sigma, (T)4 * p2 * F.psi, gi + 3 * k, nullptr, nullptr, nullptr);
    }
}

// ---- block-ordered compaction ------------------------------------------------------------------
// Every thread of the block calls this the same number of times; items flagged in one call get
// consecutive slots in thread order after all slots of earlier calls.
template <typename T>
SFX_FN int ordered_slot(bool flag, Scratch<T>& S) {
#ifdef __CUDACC__
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    int* cnt = S.cscan[S.cscan_calls & 1];     // read before the barrier by every thread: uniform
    if (lane == 0) cnt[warp] = __popc(m);
    __syncthreads();
    int before = 0, all = 0;
    for (int w = 0; w < nw; ++w) {
        const int c = cnt[w];
        before += w < warp ? c : 0;
        all += c;
    }
    const int slot = S.cscan_total + before + __popc(m & ((1u << lane) - 1u));
    __syncthreads();
    if (threadIdx.x == 0) {
        S.cscan_total += all;
        S.cscan_calls += 1;
    }
    __syncthreads();
    return flag ? slot : -1;
#else
    const int slot = S.cscan_total;
    if (flag) S.cscan_total += 1;
    return flag ? slot : -1;
#endif
}

template <typename T>
SFX_FN void face_corners(const ModelView<T>& M, const T* vert, int f, T* tri, int* ids) {
    for (int c = 0; c < 3; ++c) {
        const int v = M.faces[3 * f + c];
        ids[c] = v;
        for (int k = 0; k < 3; ++k) tri[3 * c + k] = vert[3 * v + k];
    }
}
template <typename T>
SFX_FN void tri_box(const T* tri, T* box) {
    for (int k = 0; k < 3; ++k) {
        T lo = tri[k], hi = tri[k];
        for (int c = 1; c < 3; ++c) {
            const T v = tri[3 * c + k];
            lo = v < lo ? v : lo;
            hi = v > hi ? v : hi;
        }
        box[k] = lo;
        box[3 + k] = hi;
    }
}
template <typename T>
SFX_FN bool boxes_overlap(const T* a, const T* b) {
    return a[0] <= b[3] && b[0] <= a[3] && a[1] <= b[4] && b[1] <= a[4] && a[2] <= b[5] && b[2] <= a[5];
}

// layout of the shared work area
template <typename T>
struct CollArea {
    int cap;                     // candidate capacity (a power of two)
    int cap_lean;                // ... when the packed entries move out to global memory
    unsigned char* sort_base;    // start of the sweep arrays inside the work area
    unsigned short* skey;        // [cap] extent class << 14 | box minimum along the sweep axis on a 16384-level
                                 // grid: after the sort one ascending run per class
    unsigned short* sface;       // [cap] face of the candidate
    unsigned long long* pk;      // [cap] packed: bytes 0-2 box minimum on a 256-level grid over the body's box
                                 // (rounded outwards), byte 3 part,
                                 // bytes 4-6 box maximum, byte 7 unused
    int* cseg;                   // [SFX_COLL_CLASSES + 1] segments of the sorted order, one per extent class
    unsigned int* hit;           // [(F + 31) / 32] faces that collided
    unsigned int* vtouch;        // [(V + 31) / 32] vertices of such faces
    T* part_loss;                // [SFX_NT] per-thread partial sums
    T* pbox;                     // [SFX_NPART_MAX][6]
    T* cbox;                     // [SFX_NCLUSTER_MAX][6] boxes of the face clusters
    unsigned long long* pmask;   // [SFX_NPART_MAX] admissible partner parts whose boxes overlap
};

// sweep arrays of `cap` candidates at p; the packed entries either in front of them or at q
template <typename T>
SFX_FN void coll_sort_arrays(CollArea<T>& A, unsigned char* p, int cap, unsigned char* q = nullptr) {
    A.cap = cap;
    if (q) {
        A.pk = reinterpret_cast<unsigned long long*>(q);
    } else {
        A.pk = reinterpret_cast<unsigned long long*>(p);
        p += 8L * cap;
    }
    A.skey = reinterpret_cast<unsigned short*>(p);
    A.sface = reinterpret_cast<unsigned short*>(p + 2L * cap);
}

template <typename T>
SFX_FN CollArea<T> coll_area(const ModelView<T>& M, const CollWS<T>& W, int nthreads) {
    CollArea<T> A;
    unsigned char* p = W.work;
    A.pmask = reinterpret_cast<unsigned long long*>(p); p += SFX_NPART_MAX * 8;
    A.pbox = reinterpret_cast<T*>(p); p += SFX_NPART_MAX * 6 * sizeof(T);
    A.cbox = reinterpret_cast<T*>(p); p += (size_t)M.n_clusters * 6 * sizeof(T);
    A.part_loss = reinterpret_cast<T*>(p); p += (size_t)nthreads * sizeof(T);
    A.hit = reinterpret_cast<unsigned int*>(p); p += ((M.F + 31) / 32 + 1) / 2 * 8;
    A.vtouch = reinterpret_cast<unsigned int*>(p); p += ((M.V + 31) / 32 + 1) / 2 * 8;
    A.cseg = reinterpret_cast<int*>(p); p += 8 * 4;
    const long left = (long)W.work_bytes - (long)(p - W.work);
    int cap = 0;
    for (int c = 64; (long)SFX_COLL_ENTRY * c <= left && c <= SFX_COLL_SORT_G; c <<= 1) cap = c;
    coll_sort_arrays(A, p, cap);
    A.sort_base = p;
    A.cap_lean = 0;
    for (int c = 64; 4L * c <= left && c <= SFX_COLL_SORT_G; c <<= 1) A.cap_lean = c;
    return A;
}

// Full-mesh skinning through the per-vertex weight lists: vert_g = T(v) . [vp_g; 1]
template <typename T>
SFX_FN void coll_skin_mesh(const ModelView<T>& M, Scratch<T>& S, const CollWS<T>& W) {
    SFX_FOR(v, M.V) {
        T Tm[12];
        for (int k = 0; k < 12; ++k) Tm[k] = 0;
        for (int e = M.sk_ptr[v]; e < M.sk_ptr[v + 1]; ++e) {
            const T wj = M.sk_w[e];
            const T* A = S.A + 12 * M.sk_j[e];
            for (int k = 0; k < 12; ++k) Tm[k] += wj * A[k];
        }
        const T x = W.vp_g[3 * v], y = W.vp_g[3 * v + 1], z = W.vp_g[3 * v + 2];
        for (int r = 0; r < 3; ++r)
            W.vert_g[3 * v + r] = Tm[4 * r] * x + Tm[4 * r + 1] * y + Tm[4 * r + 2] * z + Tm[4 * r + 3];
    }
    SFX_SYNC();
}

// packed entry helpers
SFX_FN unsigned int pk_lo(unsigned long long e) { return (unsigned int)e; }
SFX_FN unsigned int pk_hi(unsigned long long e) { return (unsigned int)(e >> 32); }
// do the quantised boxes of two entries overlap?  (bytes 0-2 of lo = minimum, of hi = maximum)
SFX_FN bool pk_overlap(unsigned long long a, unsigned long long b) {
    const unsigned int alo = pk_lo(a) & 0xffffffu, ahi = pk_hi(a) & 0xffffffu;
    const unsigned int blo = pk_lo(b) & 0xffffffu, bhi = pk_hi(b) & 0xffffffu;
#ifdef __CUDACC__
    // per-byte unsigned compares: all of alo <= bhi and blo <= ahi
    return (__vcmpleu4(alo, bhi) & __vcmpleu4(blo, ahi) & 0xffffffu) == 0xffffffu;
#else
    for (int k = 0; k < 3; ++k) {
        const unsigned int s = 8 * k;
        if (((alo >> s) & 255u) > ((bhi >> s) & 255u) || ((blo >> s) & 255u) > ((ahi >> s) & 255u)) return false;
    }
    return true;
#endif
}

// First index in [lo, hi) of the ascending 14-bit keys that is >= target (hi if none), found by
// the whole warp: 32 evenly spaced probes per round instead of one, three rounds for 8192 keys.
SFX_FN int warp_lower_bound(const unsigned short* skey, int lo, int hi, int target, int lane, int LW) {
#ifdef __CUDACC__
    while (lo < hi) {
        const int step = (hi - lo + LW - 1) / LW;
        const int p = lo + lane * step;
        const bool below = p < hi && (int)(skey[p] & 0x3fffu) < target;
        const int c = __popc(__ballot_sync(0xffffffffu, below));      // sorted: a prefix of the probes
        if (c == 0) return lo;
        const int pc = lo + c * step;
        lo = lo + (c - 1) * step + 1;
        hi = pc < hi ? pc : hi;
    }
    return lo;
#else
    while (lo < hi) {
        const int m = (lo + hi) >> 1;
        if ((int)(skey[m] & 0x3fffu) < target) lo = m + 1; else hi = m;
    }
    return lo;
#endif
}

// quantised extent along the sweep axis -> class; a class's bound is the largest extent in it
SFX_FN int coll_class(int ext) { return ext <= 4 ? 0 : (ext <= 16 ? 1 : (ext <= 64 ? 2 : 3)); }
SFX_FN int coll_class_bound(int c) { return c == 0 ? 4 : (c == 1 ? 16 : (c == 2 ? 64 : 255)); }

// Partners of the candidate at sorted position pos, enumerated by one warp: every lane calls
// visit(face, ok) once per step of the walk, ok marking a candidate whose part is admissible and
// whose (quantised) box overlaps; each partner shows up exactly once, in sorted order.
// The candidates are sorted by (extent class, box minimum along the sweep axis): four ascending
// runs.  In the run of class c a partner can only overlap if its key lies between the
// candidate's own key minus the class's extent bound and the candidate's maximum -- two binary
// searches, then a window whose length follows the partners' size instead of the largest
// triangle of the mesh.
template <typename T, typename VISIT>
SFX_FN void coll_partners(const CollArea<T>& A, int pos, int axis, int lane, int LW, VISIT visit) {
    const unsigned long long me = A.pk[pos];
    const unsigned long long allow = A.pmask[(pk_lo(me) >> 24) & 0x7f];
    // one 8-bit level of the box grid spans 64.25 levels of the 14-bit key grid; boxes are
    // rounded outwards, so these bounds are conservative
    const int key = (int)(A.skey[pos] & 0x3fffu);
    const int kmax = (int)(((pk_hi(me) >> (8 * axis)) & 255u) + 1u) * 65;
    for (int c = 0; c < SFX_COLL_CLASSES; ++c) {
        const int s0 = A.cseg[c], s1 = A.cseg[c + 1];
        if (s0 == s1) continue;
        const int kmin = key - (coll_class_bound(c) * 65 + 2);
        const int lo = warp_lower_bound(A.skey, s0, s1, kmin, lane, LW);       // first key >= kmin
        const int hi = warp_lower_bound(A.skey, lo, s1, kmax + 1, lane, LW);   // first key > kmax
        // Box test first (two byte-SIMD compares); the part filter and the face id only for the
        // few that pass.  The quantised boxes are rounded outwards: a superset of the exact box
        // overlaps, the separating-axis test sorts out the rest.
        for (int t0 = lo; t0 < hi; t0 += LW) {
            const int p = t0 + lane;
            bool ok = false;
            int fj = 0;
            if (p < hi) {
                const unsigned long long e = A.pk[p];
                if (pk_overlap(me, e) && p != pos) {
                    ok = (allow >> ((pk_lo(e) >> 24) & 0x7fu)) & 1ull;
                    fj = A.sface[p];
                }
            }
            visit(fj, ok);
        }
    }
}

// separating-axis test + penalty of one listed partner
template <typename T>
SFX_FN void coll_pair_loaded(int fi, const T* ti, int fj, const T* tj, T sigma, bool* hit, T* loss_i, T* gi) {
    // the package compares corner coordinates (shareVertex); so do we
    bool share = false;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            share = share || (ti[3 * a] == tj[3 * b] && ti[3 * a + 1] == tj[3 * b + 1] &&
                              ti[3 * a + 2] == tj[3 * b + 2]);
    if (!share && (fi < fj ? triangles_intersect(ti, tj) : triangles_intersect(tj, ti))) {
        *hit = true;
        pair_terms(ti, tj, sigma, loss_i, gi);
    }
}
template <typename T>
SFX_FN void coll_pair(const T* ftri, int fi, const T* ti, int fj, T sigma, bool* hit, T* loss_i, T* gi) {
    T tj[9];
    for (int d = 0; d < 9; ++d) tj[d] = ftri[(long)fj * 9 + d];
    coll_pair_loaded(fi, ti, fj, tj, sigma, hit, loss_i, gi);
}

// Search + penalty + per-vertex gradients.  On return: S.coll_loss (unweighted sum over the
// kept pairs), W.tv_g[0 .. S.n_touch) the touched vertices in ascending order and
// W.dvert_g[3 v ..] = weight * dL_pen/dvertex for each of them.
template <typename T>
SFX_FN_NOINLINE void coll_search_and_penalty(const ModelView<T>& M, Scratch<T>& S, const CollWS<T>& W,
                                             T sigma, T weight) {
    SFX_ASSUME_SHARED(M);
    SFX_ASSUME_SHARED(S);
    SFX_ASSUME_SHARED(W);
    const int F = M.F, V = M.V, NP = M.n_parts;
    CollArea<T> A = coll_area(M, W, SFX_NT);
    const T* vert = W.vert_g;
    SFX_SYNC();
    SFX_PROF_BEGIN(cb);
    // ---- boxes of all faces, then of the parts (one warp per part) ----
    SFX_FOR(f, F) {
        T tri[9], box[6];
        int ids[3];
        face_corners(M, vert, f, tri, ids);
        tri_box(tri, box);
        for (int d = 0; d < 6; ++d) W.fbox[(long)f * 6 + d] = box[d];
        for (int d = 0; d < 9; ++d) W.ftri[(long)f * 9 + d] = tri[d];
    }
    SFX_FOR(i, (F + 31) / 32) A.hit[i] = 0;
    SFX_FOR(i, (V + 31) / 32) A.vtouch[i] = 0;
    if (SFX_TID == 0) {
        S.cscan_total = 0;
        S.cscan_calls = 0;
    }
    SFX_SYNC();
    // boxes of the face clusters (one warp per cluster, two faces per lane), then of the parts
#ifdef __CUDACC__
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
        for (int c = warp; c < M.n_clusters; c += nw) {
            T lo[3] = {(T)INFINITY, (T)INFINITY, (T)INFINITY}, hi[3] = {-(T)INFINITY, -(T)INFINITY, -(T)INFINITY};
            for (int k = M.cl_ptr[c] + lane; k < M.cl_ptr[c + 1]; k += 32) {
                const T* q = W.fbox + (long)M.part_faces[k] * 6;
                for (int d = 0; d < 3; ++d) {
                    lo[d] = q[d] < lo[d] ? q[d] : lo[d];
                    hi[d] = q[3 + d] > hi[d] ? q[3 + d] : hi[d];
                }
            }
            for (int o = 16; o > 0; o >>= 1)
                for (int d = 0; d < 3; ++d) {
                    const T l2 = __shfl_xor_sync(0xffffffffu, lo[d], o), h2 = __shfl_xor_sync(0xffffffffu, hi[d], o);
                    lo[d] = l2 < lo[d] ? l2 : lo[d];
                    hi[d] = h2 > hi[d] ? h2 : hi[d];
                }
            if (lane == 0)
                for (int d = 0; d < 3; ++d) { A.cbox[6 * c + d] = lo[d]; A.cbox[6 * c + 3 + d] = hi[d]; }
        }
    }
#else
    for (int c = 0; c < M.n_clusters; ++c) {
        T lo[3] = {(T)INFINITY, (T)INFINITY, (T)INFINITY}, hi[3] = {-(T)INFINITY, -(T)INFINITY, -(T)INFINITY};
        for (int k = M.cl_ptr[c]; k < M.cl_ptr[c + 1]; ++k) {
            const T* q = W.fbox + (long)M.part_faces[k] * 6;
            for (int d = 0; d < 3; ++d) {
                lo[d] = q[d] < lo[d] ? q[d] : lo[d];
                hi[d] = q[3 + d] > hi[d] ? q[3 + d] : hi[d];
            }
        }
        for (int d = 0; d < 3; ++d) { A.cbox[6 * c + d] = lo[d]; A.cbox[6 * c + 3 + d] = hi[d]; }
    }
#endif
    SFX_SYNC();
    SFX_FOR(p, NP) {
        T lo[3] = {(T)INFINITY, (T)INFINITY, (T)INFINITY}, hi[3] = {-(T)INFINITY, -(T)INFINITY, -(T)INFINITY};
        for (int c = M.part_cl_ptr[p]; c < M.part_cl_ptr[p + 1]; ++c)
            for (int d = 0; d < 3; ++d) {
                lo[d] = A.cbox[6 * c + d] < lo[d] ? A.cbox[6 * c + d] : lo[d];
                hi[d] = A.cbox[6 * c + 3 + d] > hi[d] ? A.cbox[6 * c + 3 + d] : hi[d];
            }
        for (int d = 0; d < 3; ++d) { A.pbox[6 * p + d] = lo[d]; A.pbox[6 * p + 3 + d] = hi[d]; }
    }
    SFX_SYNC();
    SFX_FOR(p, NP) {
        unsigned long long m = 0;
        const unsigned long long allow = M.part_allow[p];
        const bool empty_p = M.part_ptr[p] == M.part_ptr[p + 1];
        for (int q = 0; q < NP; ++q)
            if (((allow >> q) & 1ull) && !empty_p && M.part_ptr[q] != M.part_ptr[q + 1] &&
                boxes_overlap(A.pbox + 6 * p, A.pbox + 6 * q))
                m |= 1ull << q;
        A.pmask[p] = m;
    }
    SFX_SYNC();
    // sweep axis: the longest side of the body's box
    int axis = 0;
    {
        T best = -1;
        for (int d = 0; d < 3; ++d) {
            T lo = (T)INFINITY, hi = -(T)INFINITY;
            for (int p = 0; p < NP; ++p)
                if (M.part_ptr[p] != M.part_ptr[p + 1]) {
                    lo = A.pbox[6 * p + d] < lo ? A.pbox[6 * p + d] : lo;
                    hi = A.pbox[6 * p + 3 + d] > hi ? A.pbox[6 * p + 3 + d] : hi;
                }
            if (hi - lo > best) { best = hi - lo; axis = d; }
        }
    }
    // the body's box: origin and scale of the 256-level (boxes) and 65536-level (sweep key) grids
    T blo[3], bsc[3];
    for (int d = 0; d < 3; ++d) {
        T lo = (T)INFINITY, hi = -(T)INFINITY;
        for (int p = 0; p < NP; ++p)
            if (M.part_ptr[p] != M.part_ptr[p + 1]) {
                lo = A.pbox[6 * p + d] < lo ? A.pbox[6 * p + d] : lo;
                hi = A.pbox[6 * p + 3 + d] > hi ? A.pbox[6 * p + 3 + d] : hi;
            }
        blo[d] = lo;
        bsc[d] = hi > lo ? (T)255 / (hi - lo) : (T)0;
    }
    // ---- candidate faces: their box reaches into the box of an admissible partner part ----
    // Compacted in part order into the sweep arrays: first the shared ones; if they do not fit,
    // once more into the block's global arrays.
    for (int attempt = 0; attempt < 2; ++attempt) {
        for (int k0 = 0; k0 < F; k0 += SFX_NT) {
            const int k = k0 + SFX_TID;
            bool cand = false;
            int f = 0, part = 0;
            int key = 0;
            if (k < F) {
                f = M.part_faces[k];
                part = M.face_part[f];
                const unsigned long long m = A.pmask[part];
                if (m) {
                    T box[6];
                    for (int d = 0; d < 6; ++d) box[d] = W.fbox[(long)f * 6 + d];
                    // ... the box of one of the partner part's clusters, to be precise
                    for (int q = 0; q < NP && !cand; ++q)
                        if (((m >> q) & 1ull) && boxes_overlap(box, A.pbox + 6 * q))
                            for (int c = M.part_cl_ptr[q]; c < M.part_cl_ptr[q + 1] && !cand; ++c)
                                cand = boxes_overlap(box, A.cbox + 6 * c);
                    // 14-bit key of the box minimum and the extent class, both on grids derived
                    // from the same 256-level box grid as the packed entries below
                    const float qlo = (float)((box[axis] - blo[axis]) * bsc[axis]);
                    const float qhi = (float)((box[3 + axis] - blo[axis]) * bsc[axis]);
                    int k14 = (int)floorf(qlo * 64.25f);
                    k14 = k14 < 0 ? 0 : (k14 > 16383 ? 16383 : k14);
                    int e0 = (int)floorf(qlo - 1e-3f), e1 = (int)ceilf(qhi + 1e-3f);
                    e0 = e0 < 0 ? 0 : (e0 > 255 ? 255 : e0);
                    e1 = e1 < 0 ? 0 : (e1 > 255 ? 255 : e1);
                    key = (coll_class(e1 - e0) << 14) | k14;
                }
            }
            const int slot = ordered_slot(cand, S);
            if (cand && slot < A.cap) {
                A.skey[slot] = (unsigned short)key;
                A.sface[slot] = (unsigned short)f;
            }
        }
        SFX_SYNC();
        if (S.cscan_total <= A.cap) break;           // uniform: written before the barrier
        if (attempt == 0 && W.sort_g != nullptr && A.cap < SFX_COLL_SORT_G) {
            // second tier: keys / faces / parts stay in shared memory, the quantised boxes move
            // to the block's global arrays; third tier: everything in global memory
            if (S.cscan_total <= A.cap_lean) coll_sort_arrays(A, A.sort_base, A.cap_lean, W.sort_g);
            else coll_sort_arrays(A, W.sort_g, SFX_COLL_SORT_G);
            SFX_SYNC();
            if (SFX_TID == 0) {
                S.cscan_total = 0;
                S.cscan_calls = 0;
            }
            SFX_SYNC();
        } else {
            if (SFX_TID == 0) S.coll_overflow = 1;
            break;
        }
    }
    const int ncand = S.cscan_total < A.cap ? S.cscan_total : A.cap;
    if (SFX_TID == 0 && ncand > S.coll_max_cand) S.coll_max_cand = ncand;
    // ---- sort by (key, face): bitonic network over the next power of two ----
    int n2 = 1;
    while (n2 < ncand) n2 <<= 1;
    SFX_FOR(i, n2 - ncand) {
        A.skey[ncand + i] = 0xffff;
        A.sface[ncand + i] = 0xffff;
    }
    SFX_SYNC();
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            SFX_FOR(t, n2) {
                const int u = t ^ j;
                if (u > t) {
                    const unsigned short kt = A.skey[t], ku = A.skey[u];
                    const unsigned short ft = A.sface[t], fu = A.sface[u];
                    const bool greater = (((unsigned)kt << 16) | ft) > (((unsigned)ku << 16) | fu);
                    if (greater == ((t & k) == 0)) {
                        A.skey[t] = ku; A.skey[u] = kt;
                        A.sface[t] = fu; A.sface[u] = ft;
                    }
                }
            }
            SFX_SYNC();
        }
    // ---- quantised boxes (256 levels over the body's box, rounded outwards) ----
    SFX_FOR(c, ncand) {
        const T* fb = W.fbox + (long)A.sface[c] * 6;
        int q[6];
        for (int d = 0; d < 3; ++d) {
            const float lo = (float)((fb[d] - blo[d]) * bsc[d]) - 1e-3f;
            const float hi = (float)((fb[3 + d] - blo[d]) * bsc[d]) + 1e-3f;
            int a = (int)floorf(lo), b = (int)ceilf(hi);
            a = a < 0 ? 0 : (a > 255 ? 255 : a);
            b = b < 0 ? 0 : (b > 255 ? 255 : b);
            q[d] = a;
            q[3 + d] = b;
        }
        A.pk[c] = (unsigned long long)((unsigned)q[0] | ((unsigned)q[1] << 8) | ((unsigned)q[2] << 16) |
                                       ((unsigned)M.face_part[A.sface[c]] << 24)) |
                  ((unsigned long long)((unsigned)q[3] | ((unsigned)q[4] << 8) | ((unsigned)q[5] << 16)) << 32);
    }
    // where each extent class starts in the sorted order
    SFX_FOR(c, SFX_COLL_CLASSES + 1) {
        int a = 0, b = ncand;
        while (a < b) { const int m = (a + b) >> 1; if ((int)(A.skey[m] >> 14) < c) a = m + 1; else b = m; }
        A.cseg[c] = a;
    }
    SFX_SYNC();
    // ---- walk + narrow phase, a warp per candidate, in chunks ----
    // Walk: the warp lists the box-overlapping, admissible partners of the candidates it owns
    // (sorted positions warp, warp + 16, ...) in its own region, (count, faces...) per candidate;
    // a candidate is only started while half the region is free, and when the region cannot take
    // the next one the warp works the list off before it walks on.
    // Narrow phase: separating-axis test, penalty of the own cone and gradient w.r.t. the own
    // corners (a gather: no atomics on values, a fixed summation order).
#ifdef __CUDACC__
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NWP = blockDim.x >> 5, LW = 32;
#else
    const int lane = 0, warp = 0, NWP = 1, LW = 1;
#endif
    const int HCW = W.hits_cap * LW;                 // the warp's region: its lanes' regions together
    unsigned short* myhits = W.hits_g + (long)warp * HCW;
    T my_loss = 0;
    int walk_iters = 0, walk_hits = 0;      // diagnostics (warp 0): window steps / 32, listed partners
    SFX_PROF_END(S, 9, cb);
    for (int pos0 = warp; pos0 < ncand;) {
        // ---- walk: a warp per candidate, lanes side by side through the window ----
        int w = 0, pos1 = pos0;
        SFX_PROF_BEGIN(cw);
        for (; pos1 < ncand && w + HCW / 2 <= HCW; pos1 += NWP) {
            const int head = w++;
            int cnt = 0;
            coll_partners(A, pos1, axis, lane, LW, [&](int fj, bool ok) {
#ifdef __CUDACC__
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                const int slot = w + __popc(m & ((1u << lane) - 1u)), n = __popc(m);
#else
                const int slot = w, n = ok ? 1 : 0;
#endif
                walk_iters += 1;
                if (ok && slot < HCW) myhits[slot] = (unsigned short)fj;
                w += n;
                cnt += n;
            });
            // a candidate with more box partners than the region holds (or than a 16-bit count)
            // is walked again in the narrow phase, partner by partner
            walk_hits += cnt;
            if (w > HCW || cnt >= 0xffff) { w = head + 1; cnt = 0xffff; }
            if (lane == 0) myhits[head] = (unsigned short)cnt;
        }
        SFX_SYNCWARP();
        SFX_PROF_END(S, 13, cw);
        SFX_PROF_BEGIN(ch);
        // ---- narrow phase: the lanes share a candidate's partners; per-lane partial sums,
        // combined in a fixed order ----
        int r = 0;
        for (int pos = pos0; pos < pos1; pos += NWP) {
            const int cnt = myhits[r++];
            if (cnt == 0) continue;
            const int fi = A.sface[pos];
            T ti[9], acc[10];
            int idi[3];
            face_corners(M, vert, fi, ti, idi);
            for (int d = 0; d < 10; ++d) acc[d] = 0;
            bool hit = false;
            if (cnt == 0xffff) {
                coll_partners(A, pos, axis, lane, LW, [&](int fj, bool ok) {
                    if (ok) coll_pair(W.ftri, fi, ti, fj, sigma, &hit, acc + 9, acc);
                });
            } else {
                // the partner's corners of the next round travel while this round's test runs
                T tn[9];
                int fn = 0;
                if (lane < cnt) {
                    fn = myhits[r + lane];
                    for (int d = 0; d < 9; ++d) tn[d] = W.ftri[(long)fn * 9 + d];
                }
                for (int h = lane; h < cnt; h += LW) {
                    T tj[9];
                    const int fj = fn;
                    for (int d = 0; d < 9; ++d) tj[d] = tn[d];
                    if (h + LW < cnt) {
                        fn = myhits[r + h + LW];
                        for (int d = 0; d < 9; ++d) tn[d] = W.ftri[(long)fn * 9 + d];
                    }
                    coll_pair_loaded(fi, ti, fj, tj, sigma, &hit, acc + 9, acc);
                }
                r += cnt;
            }
#ifdef __CUDACC__
            hit = __any_sync(0xffffffffu, hit);
            if (hit)
                for (int o = 16; o > 0; o >>= 1)
                    for (int d = 0; d < 10; ++d) acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], o);
#endif
            if (hit && lane == 0) {
                my_loss += acc[9];
                for (int d = 0; d < 9; ++d) W.dtri_g[(long)fi * 9 + d] = acc[d];
#ifdef __CUDACC__
                atomicOr(&A.hit[fi >> 5], 1u << (fi & 31));
                for (int a = 0; a < 3; ++a) atomicOr(&A.vtouch[idi[a] >> 5], 1u << (idi[a] & 31));
#else
                A.hit[fi >> 5] |= 1u << (fi & 31);
                for (int a = 0; a < 3; ++a) A.vtouch[idi[a] >> 5] |= 1u << (idi[a] & 31);
#endif
            }
        }
        SFX_SYNCWARP();
        SFX_PROF_END(S, 14, ch);
        pos0 = pos1;
    }
    if (SFX_TID == 0) {
        if (walk_iters > S.coll_max_iters) S.coll_max_iters = walk_iters;
        if (walk_hits > S.coll_max_hits) S.coll_max_hits = walk_hits;
    }
    SFX_PROF_BEGIN(cn);
    A.part_loss[SFX_TID] = my_loss;
    {
        const T* pl = A.part_loss;
        const T tot = block_reduce<T>(SFX_NT, [=](int i) { return pl[i]; }, OpAdd<T>(), (T)0, &S.red[9]);
        if (SFX_TID == 0) S.coll_loss = tot;
    }
    SFX_PROF_END(S, 10, cn);
    SFX_PROF_BEGIN(cg);
    // ---- per-vertex gradients through the static vertex -> face table; touched list ----
    if (SFX_TID == 0) {
        S.cscan_total = 0;
        S.cscan_calls = 0;
    }
    SFX_SYNC();
    for (int v0 = 0; v0 < V; v0 += SFX_NT) {
        const int v = v0 + SFX_TID;
        bool touched = false;
        if (v < V && ((A.vtouch[v >> 5] >> (v & 31)) & 1u)) {
            T acc[3] = {0, 0, 0};
            for (int e = M.vf_ptr[v]; e < M.vf_ptr[v + 1]; ++e) {
                const int fc = M.vf_idx[e], f = fc >> 2, c = fc & 3;
                if ((A.hit[f >> 5] >> (f & 31)) & 1u)
                    for (int k = 0; k < 3; ++k) acc[k] += W.dtri_g[(long)f * 9 + 3 * c + k];
            }
            for (int k = 0; k < 3; ++k) W.dvert_g[3 * v + k] = weight * acc[k];
            touched = true;
        }
        const int slot = ordered_slot(touched, S);
        if (touched) W.tv_g[slot] = (unsigned short)v;
    }
    SFX_SYNC();
    if (SFX_TID == 0) {
        S.n_touch = S.cscan_total;
        if (S.cscan_total > S.coll_max_touch) S.coll_max_touch = S.cscan_total;
    }
    SFX_PROF_END(S, 11, cg);
#ifdef __CUDACC__
    // the work area goes back to the blend ring, which TMA (async proxy) writes next
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
    SFX_SYNC();
}

// Adjoint of the skinning of the touched vertices: S.dA += ..., vert_g[3 t ..] <- dL/dv_posed of
// touched vertex t (compact order; the posed vertices are no longer needed at this point).
template <typename T>
SFX_FN_NOINLINE void coll_skin_adjoint(const ModelView<T>& M, Scratch<T>& S, const CollWS<T>& W) {
    SFX_ASSUME_SHARED(M);
    SFX_ASSUME_SHARED(S);
    SFX_ASSUME_SHARED(W);
    const int nt = S.n_touch;
    SFX_PROF_BEGIN(ca);
    SFX_SYNC();
    SFX_FOR(i, SFX_NJ * 12) {
        const int j = i / 12, r = (i % 12) / 4, cc = i % 4;
        T acc = 0;
        for (int t = 0; t < nt; ++t) {
            const int v = W.tv_g[t];
            for (int e = M.sk_ptr[v]; e < M.sk_ptr[v + 1]; ++e)
                if (M.sk_j[e] == j)
                    acc += M.sk_w[e] * W.dvert_g[3 * v + r] * (cc < 3 ? W.vp_g[3 * v + cc] : (T)1);
        }
        S.dA[i] += acc;
    }
    SFX_SYNC();
    SFX_FOR(t, nt) {
        const int v = W.tv_g[t];
        T R[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int e = M.sk_ptr[v]; e < M.sk_ptr[v + 1]; ++e) {
            const T wj = M.sk_w[e];
            const T* A = S.A + 12 * M.sk_j[e];
            for (int r = 0; r < 3; ++r)
                for (int k = 0; k < 3; ++k) R[3 * r + k] += wj * A[4 * r + k];
        }
        const T d0 = W.dvert_g[3 * v], d1 = W.dvert_g[3 * v + 1], d2 = W.dvert_g[3 * v + 2];
        for (int k = 0; k < 3; ++k) W.vert_g[3 * t + k] = R[k] * d0 + R[3 + k] * d1 + R[6 + k] * d2;
    }
    SFX_SYNC();
    SFX_PROF_END(S, 12, ca);
}

}  // namespace sfx
