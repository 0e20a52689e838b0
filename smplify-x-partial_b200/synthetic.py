"""Seeded synthetic SMPL-X-schema assets.

The licensed ``SMPLX_{NEUTRAL,MALE,FEMALE}.npz`` files, the VPoser v1 weights
and ``gmm_08.pkl`` cannot be shipped or downloaded.  This
module builds stand-ins with the *same key set, shapes and dtypes* as the real
files so every code path (loader, kernels, oracle) is exercised on data of the
real size: V = 10475 vertices, F = 20908 faces, J = 55 joints, 400 shape
directions (300 identity + 100 expression), 486 pose-corrective directions,
45x45 hand PCA bases, 51 static + 79x17 dynamic face-landmark tables.

Geometry is a procedural tube-man: every bone of the SMPL-X kinematic tree
(tree recovered in SURVEY.md section 3.5 from ``smplifyx/smplx_parts_segm.pkl``)
gets a tube of vertices with smooth skinning weights, so that 2-D keypoints of
a real person (the ``demo/`` frames) are a meaningful fitting target.  A real
model file loads through the same loader unchanged (``body_model.load_model``).

Pure numpy + ``numpy.random.default_rng(seed)``: the same seed gives the same
bytes on every box.
"""
import numpy as np

NUM_VERTS = 10475
NUM_FACES = 20908
NUM_JOINTS = 55
NUM_POSE_BASIS = 486
NUM_SHAPE_TOTAL = 400
EXPR_OFFSET = 300

# Kinematic tree of SMPL-X (parent of joint i), SURVEY.md section 3.5.
PARENTS = np.array(
    [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19,
     15, 15, 15, 20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35, 20, 37, 38,
     21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50, 21, 52, 53],
    dtype=np.int64)

# Vertex ids that smplx.vertex_ids['smplx'] selects as extra joints
# (nose, eyes, ears, feet, finger tips); order = VertexJointSelector order.
EXTRA_VERTEX_IDS = np.array(
    [9120, 9929, 9448, 616, 6,                      # nose reye leye rear lear
     5770, 5780, 8846, 8463, 8474, 8635,            # LBigToe LSmallToe LHeel RBigToe RSmallToe RHeel
     5361, 4933, 5058, 5169, 5286,                  # left thumb index middle ring pinky
     8079, 7669, 7794, 7905, 8022],                 # right thumb index middle ring pinky
    dtype=np.int64)


def _rest_joints():
    """Approximate T-pose joint locations of an adult (metres, y up, +x = subject's left)."""
    J = np.zeros((NUM_JOINTS, 3), dtype=np.float64)
    J[0] = (0.00, -0.35, 0.00)      # pelvis
    J[1] = (0.07, -0.44, 0.00)      # left hip
    J[2] = (-0.07, -0.44, 0.00)
    J[3] = (0.00, -0.23, -0.02)     # spine1
    J[4] = (0.10, -0.83, 0.00)      # knees
    J[5] = (-0.10, -0.83, 0.00)
    J[6] = (0.00, -0.09, 0.00)      # spine2
    J[7] = (0.09, -1.24, -0.03)     # ankles
    J[8] = (-0.09, -1.24, -0.03)
    J[9] = (0.00, -0.03, 0.00)      # spine3
    J[10] = (0.11, -1.30, 0.09)     # feet
    J[11] = (-0.11, -1.30, 0.09)
    J[12] = (0.00, 0.18, -0.03)     # neck
    J[13] = (0.06, 0.09, -0.02)     # collars
    J[14] = (-0.06, 0.09, -0.02)
    J[15] = (0.00, 0.27, 0.01)      # head
    J[16] = (0.18, 0.11, -0.02)     # shoulders
    J[17] = (-0.18, 0.11, -0.02)
    J[18] = (0.44, 0.10, -0.03)     # elbows
    J[19] = (-0.44, 0.10, -0.03)
    J[20] = (0.69, 0.10, -0.02)     # wrists
    J[21] = (-0.69, 0.10, -0.02)
    J[22] = (0.00, 0.26, 0.03)      # jaw
    J[23] = (0.032, 0.335, 0.075)   # eyes
    J[24] = (-0.032, 0.335, 0.075)
    # fingers: index, middle, pinky, ring, thumb  (3 joints each)
    finger_z = [0.025, 0.002, -0.040, -0.020]
    finger_len = [(0.095, 0.030, 0.022), (0.097, 0.032, 0.024),
                  (0.085, 0.020, 0.016), (0.092, 0.028, 0.022)]
    for side, base, sgn in ((0, 25, 1.0), (1, 40, -1.0)):
        wrist = J[20 + side]
        for f in range(4):
            x = wrist[0]
            for k in range(3):
                x = x + sgn * finger_len[f][k]
                J[base + 3 * f + k] = (x, wrist[1] - 0.002 * k, wrist[2] + finger_z[f])
        J[base + 12] = (wrist[0] + sgn * 0.025, wrist[1] - 0.010, wrist[2] + 0.030)
        J[base + 13] = (wrist[0] + sgn * 0.050, wrist[1] - 0.015, wrist[2] + 0.052)
        J[base + 14] = (wrist[0] + sgn * 0.075, wrist[1] - 0.020, wrist[2] + 0.066)
    return J


def _bone_radius(j):
    if j in (0, 3, 6, 9):
        return 0.13
    if j in (1, 2):
        return 0.075
    if j in (4, 5):
        return 0.05
    if j in (7, 8, 10, 11):
        return 0.035
    if j in (12,):
        return 0.055
    if j in (13, 14):
        return 0.06
    if j == 15:
        return 0.095
    if j in (16, 17):
        return 0.045
    if j in (18, 19):
        return 0.035
    if j in (20, 21):
        return 0.03
    if j == 22:
        return 0.04
    if j in (23, 24):
        return 0.012
    return 0.008            # fingers


def _children(parents):
    ch = [[] for _ in range(len(parents))]
    for i, p in enumerate(parents):
        if p >= 0:
            ch[p].append(i)
    return ch


def make_smplx_like(seed=0, dtype=np.float32, full_shape_space=True):
    """Returns a dict with the SMPL-X npz key set (see module docstring)."""
    rng = np.random.default_rng(seed)
    J = _rest_joints()
    parents = PARENTS
    children = _children(parents)

    # ---- bone segments: joint -> tip --------------------------------------------
    tips = np.zeros_like(J)
    for j in range(NUM_JOINTS):
        if children[j]:
            # follow the "main" child so the spine / limbs form continuous tubes
            main = {0: 3, 9: 12, 12: 15, 15: 15, 20: 31, 21: 46}.get(j, children[j][0])
            tips[j] = J[main] if main != j else J[j] + np.array([0, 0.12, 0.0])
        else:
            d = J[j] - J[parents[j]]
            n = np.linalg.norm(d)
            tips[j] = J[j] + d / max(n, 1e-6) * max(0.6 * n, 0.015)
    tips[15] = J[15] + np.array([0.0, 0.13, 0.01])   # head extends up
    tips[22] = J[22] + np.array([0.0, -0.05, 0.06])  # jaw extends to the chin

    # ---- vertex budget per bone -------------------------------------------------
    seg_len = np.linalg.norm(tips - J, axis=1)
    radius = np.array([_bone_radius(j) for j in range(NUM_JOINTS)])
    area = (seg_len + radius) * radius
    area[15] *= 6.0          # dense head/face like the real mesh
    area[22] *= 3.0
    area[25:] *= 3.0         # dense hands
    n_special = 21 + 51 * 3 + 128 * 3      # extra joints + static landmark tris + contour strip
    budget = NUM_VERTS - n_special
    counts = np.maximum(24, np.floor(area / area.sum() * budget)).astype(np.int64)
    counts[0] += budget - counts.sum()
    assert counts.min() > 0 and counts.sum() == budget

    verts = np.zeros((NUM_VERTS, 3), dtype=np.float64)
    owner = np.zeros(NUM_VERTS, dtype=np.int64)       # bone of each vertex
    upar = np.zeros(NUM_VERTS, dtype=np.float64)      # position along the bone in [0, 1]
    faces = []
    perm = rng.permutation(NUM_VERTS)
    # keep the fixed extra-joint ids and a reserved block out of the tube pool
    reserved = set(int(v) for v in EXTRA_VERTEX_IDS)
    pool = [int(v) for v in perm if int(v) not in reserved]
    cursor = 0

    def take(n):
        nonlocal cursor
        out = pool[cursor:cursor + n]
        cursor += n
        assert len(out) == n
        return np.array(out, dtype=np.int64)

    for j in range(NUM_JOINTS):
        n = int(counts[j])
        segs = int(np.clip(round(np.sqrt(n * 2.0 * np.pi * radius[j] /
                                         max(seg_len[j], 1e-3))), 4, 48))
        rings = n // segs
        ids = take(rings * segs)
        extra = take(n - rings * segs)
        axis = tips[j] - J[j]
        L = np.linalg.norm(axis)
        a = axis / max(L, 1e-9)
        e1 = np.cross(a, np.array([0.0, 0.0, 1.0]))
        if np.linalg.norm(e1) < 1e-3:
            e1 = np.cross(a, np.array([1.0, 0.0, 0.0]))
        e1 /= np.linalg.norm(e1)
        e2 = np.cross(a, e1)
        for r in range(rings):
            u = (r + 0.5) / rings
            # ellipsoidal profile so tubes close at the ends
            prof = radius[j] * (0.55 + 0.45 * np.sin(np.pi * u))
            for s in range(segs):
                phi = 2.0 * np.pi * (s + 0.5 * (r % 2)) / segs
                vid = ids[r * segs + s]
                verts[vid] = (J[j] + a * (u * L) +
                              prof * (np.cos(phi) * e1 + np.sin(phi) * e2))
                owner[vid] = j
                upar[vid] = u
        for r in range(rings - 1):
            for s in range(segs):
                v00 = ids[r * segs + s]
                v01 = ids[r * segs + (s + 1) % segs]
                v10 = ids[(r + 1) * segs + s]
                v11 = ids[(r + 1) * segs + (s + 1) % segs]
                faces.append((v00, v01, v11))
                faces.append((v00, v11, v10))
        for vid in extra:      # leftovers: jitter inside the tube
            u = rng.uniform(0.1, 0.9)
            phi = rng.uniform(0, 2 * np.pi)
            verts[vid] = (J[j] + a * (u * L) + 0.9 * radius[j] *
                          (np.cos(phi) * e1 + np.sin(phi) * e2))
            owner[vid] = j
            upar[vid] = u

    # ---- extra-joint vertices at anatomical spots -------------------------------
    head = J[15]
    spots = {
        9120: (head + (0.0, 0.045, 0.125), 15),       # nose
        9929: (head + (-0.032, 0.065, 0.095), 15),    # right eye
        9448: (head + (0.032, 0.065, 0.095), 15),     # left eye
        616: (head + (-0.078, 0.05, 0.0), 15),        # right ear
        6: (head + (0.078, 0.05, 0.0), 15),           # left ear
        5770: (J[10] + (0.02, -0.02, 0.09), 10),      # L big toe
        5780: (J[10] + (0.06, -0.02, 0.06), 10),      # L small toe
        8846: (J[7] + (0.0, -0.06, -0.05), 7),        # L heel
        8463: (J[11] + (-0.02, -0.02, 0.09), 11),
        8474: (J[11] + (-0.06, -0.02, 0.06), 11),
        8635: (J[8] + (0.0, -0.06, -0.05), 8),
    }
    tip_joints_l = [39, 27, 30, 36, 33]      # thumb index middle ring pinky (3rd joint)
    tip_joints_r = [54, 42, 45, 51, 48]
    for k, vid in enumerate(EXTRA_VERTEX_IDS[11:16]):
        j = tip_joints_l[k]
        spots[int(vid)] = (tips[j], j)
    for k, vid in enumerate(EXTRA_VERTEX_IDS[16:21]):
        j = tip_joints_r[k]
        spots[int(vid)] = (tips[j], j)
    for vid, (p, j) in spots.items():
        verts[vid] = p
        owner[vid] = j
        upar[vid] = 0.9

    # ---- static face landmarks: 51 small triangles on the face ------------------
    lmk_pts = []
    for i in range(5):       # right brow
        lmk_pts.append((-0.060 + 0.010 * i, 0.085 + 0.004 * np.sin(i), 0.098))
    for i in range(5):       # left brow
        lmk_pts.append((0.020 + 0.010 * i, 0.085 + 0.004 * np.sin(i), 0.098))
    for i in range(4):       # nose bridge
        lmk_pts.append((0.0, 0.072 - 0.009 * i, 0.105 + 0.006 * i))
    for i in range(5):       # nostrils
        lmk_pts.append((-0.016 + 0.008 * i, 0.032, 0.108))
    for i in range(6):       # right eye
        a = 2 * np.pi * i / 6
        lmk_pts.append((-0.032 + 0.012 * np.cos(a), 0.065 + 0.005 * np.sin(a), 0.094))
    for i in range(6):       # left eye
        a = 2 * np.pi * i / 6
        lmk_pts.append((0.032 + 0.012 * np.cos(a), 0.065 + 0.005 * np.sin(a), 0.094))
    for i in range(12):      # outer mouth
        a = 2 * np.pi * i / 12
        lmk_pts.append((0.026 * np.cos(a), 0.005 + 0.011 * np.sin(a), 0.100))
    for i in range(8):       # inner mouth
        a = 2 * np.pi * i / 8
        lmk_pts.append((0.015 * np.cos(a), 0.005 + 0.005 * np.sin(a), 0.098))
    lmk_pts = np.array(lmk_pts, dtype=np.float64)
    assert lmk_pts.shape[0] == 51
    lmk_faces_idx = np.zeros(51, dtype=np.int64)
    lmk_bary = np.zeros((51, 3), dtype=np.float64)
    for i in range(51):
        ids = take(3)
        c = head + lmk_pts[i]
        tri = c + 0.004 * rng.normal(size=(3, 3)) * np.array([1.0, 1.0, 0.2])
        verts[ids] = tri
        mouth = i >= 31
        owner[ids] = 22 if mouth else 15
        upar[ids] = 0.5
        lmk_faces_idx[i] = len(faces)
        faces.append(tuple(ids))
        b = rng.dirichlet(np.ones(3) * 4.0)
        lmk_bary[i] = b

    # ---- dynamic contour: strip of 128 triangles around the jaw line ------------
    strip_faces = np.zeros(128, dtype=np.int64)
    for k in range(128):
        ang = np.deg2rad(-127.0 + 2.0 * k)          # -127 .. +127 degrees around the head
        ids = take(3)
        c = head + np.array([0.085 * np.sin(ang), 0.01 - 0.035 * np.cos(ang / 2.0) ** 2,
                             0.085 * np.cos(ang) + 0.01])
        verts[ids] = c + 0.004 * rng.normal(size=(3, 3))
        owner[ids] = 22 if abs(np.rad2deg(ang)) < 50 else 15
        upar[ids] = 0.5
        strip_faces[k] = len(faces)
        faces.append(tuple(ids))
    dyn_faces = np.zeros((79, 17), dtype=np.int64)
    dyn_bary = np.zeros((79, 17, 3), dtype=np.float64)
    for y in range(79):
        # LUT row index as produced by find_dynamic_lmk_idx_and_bcoords: 0..39 = yaw 0..39 deg,
        # 40..78 = yaw -1..-39 deg.
        yaw = y if y <= 39 else -(y - 39)
        for i in range(17):
            ang = -80.0 + 10.0 * i + 0.8 * yaw        # contour slides with head yaw
            k = int(np.clip(round((ang + 127.0) / 2.0), 0, 127))
            dyn_faces[y, i] = strip_faces[k]
            dyn_bary[y, i] = rng.dirichlet(np.ones(3) * 4.0)
    assert cursor == len(pool), (cursor, len(pool))

    # ---- pad / trim the face list to the SMPL-X count ---------------------------
    faces = [tuple(int(x) for x in f) for f in faces]
    while len(faces) < NUM_FACES:
        # bridging triangles between random vertices of neighbouring bones
        j = int(rng.integers(1, NUM_JOINTS))
        a = np.flatnonzero(owner == j)
        b = np.flatnonzero(owner == parents[j])
        faces.append((int(rng.choice(a)), int(rng.choice(a)), int(rng.choice(b))))
    assert len(faces) >= NUM_FACES
    protected = set(int(f) for f in lmk_faces_idx) | set(int(f) for f in strip_faces)
    if len(faces) > NUM_FACES:
        keep = [i for i in range(len(faces))]
        drop = [i for i in keep if i not in protected][NUM_FACES - len(faces):]
        drop = set(drop)
        remap = {}
        new_faces = []
        for i, f in enumerate(faces):
            if i in drop:
                continue
            remap[i] = len(new_faces)
            new_faces.append(f)
        faces = new_faces
        lmk_faces_idx = np.array([remap[int(f)] for f in lmk_faces_idx])
        dyn_faces = np.vectorize(lambda f: remap[int(f)])(dyn_faces)
    faces = np.array(faces, dtype=np.uint32)
    assert faces.shape == (NUM_FACES, 3)

    # ---- skinning weights: blend with parent near the joint, child near the tip --
    W = np.zeros((NUM_VERTS, NUM_JOINTS), dtype=np.float64)
    for v in range(NUM_VERTS):
        j = int(owner[v])
        u = upar[v]
        W[v, j] += 1.0
        p = parents[j]
        if p >= 0:
            W[v, p] += max(0.0, 0.45 - 1.4 * u)
            if parents[p] >= 0:
                W[v, parents[p]] += max(0.0, 0.10 - 0.6 * u)
        if children[j]:
            # nearest child
            ch = children[j]
            dists = [np.linalg.norm(verts[v] - J[c]) for c in ch]
            c = ch[int(np.argmin(dists))]
            W[v, c] += max(0.0, 1.4 * u - 0.95)
    W *= (1.0 + 0.05 * rng.uniform(size=W.shape)) * (W > 0)
    W /= W.sum(axis=1, keepdims=True)

    # ---- joint regressor: each joint from its nearest vertices ------------------
    Jreg = np.zeros((NUM_JOINTS, NUM_VERTS), dtype=np.float64)
    for j in range(NUM_JOINTS):
        d = np.linalg.norm(verts - J[j], axis=1)
        nn = np.argsort(d)[:64]
        w = np.exp(-(d[nn] / (d[nn].mean() + 1e-9)) ** 2)
        Jreg[j, nn] = w / w.sum()

    # ---- shape / expression / pose-corrective directions ------------------------
    n_sh = NUM_SHAPE_TOTAL if full_shape_space else 20
    shapedirs = np.zeros((NUM_VERTS, 3, NUM_SHAPE_TOTAL), dtype=dtype)
    centred = verts - verts.mean(axis=0)
    for k in range(10):
        # smooth low-frequency deformation fields: scale, limb length, girth ...
        fr = rng.normal(size=(3, 3)) * (1.0 + 0.5 * k)
        ph = rng.uniform(0, 2 * np.pi, size=3)
        fld = np.sin(centred @ fr.T + ph) * (0.02 / (1.0 + 0.35 * k))
        fld += centred * rng.normal(size=3) * (0.03 / (1.0 + k))
        shapedirs[:, :, k] = fld.astype(dtype)
    if full_shape_space:
        shapedirs[:, :, 10:EXPR_OFFSET] = (
            rng.standard_normal(size=(NUM_VERTS, 3, EXPR_OFFSET - 10), dtype=np.float32)
            * 1e-3).astype(dtype)
    headmask = ((owner == 15) | (owner == 22)).astype(np.float64)[:, None]
    for k in range(10):
        fr = rng.normal(size=(3, 3)) * 12.0
        ph = rng.uniform(0, 2 * np.pi, size=3)
        fld = np.sin((verts - head) @ fr.T + ph) * 0.004 * headmask
        shapedirs[:, :, EXPR_OFFSET + k] = fld.astype(dtype)
    if full_shape_space:
        shapedirs[:, :, EXPR_OFFSET + 10:] = (
            rng.standard_normal(size=(NUM_VERTS, 3, NUM_SHAPE_TOTAL - EXPR_OFFSET - 10),
                                dtype=np.float32) * 5e-4 * headmask[:, :, None]).astype(dtype)

    # pose correctives: dense, small, strongest for the joints that move the vertex
    posedirs = rng.standard_normal(size=(NUM_VERTS, 3, NUM_POSE_BASIS), dtype=np.float32)
    gain = np.full((NUM_VERTS, NUM_JOINTS - 1), 4e-4, dtype=np.float32)
    for v in range(NUM_VERTS):
        j = int(owner[v])
        for jj in (j, int(parents[j])):
            if jj >= 1:
                gain[v, jj - 1] = 6e-3
    posedirs *= np.repeat(gain, 9, axis=1)[:, None, :]
    posedirs = posedirs.astype(dtype)

    # ---- hand PCA -----------------------------------------------------------------
    def _orth(n):
        q, r = np.linalg.qr(rng.normal(size=(n, n)))
        return q * np.sign(np.diag(r))
    hands_componentsl = _orth(45).astype(dtype)
    hands_componentsr = _orth(45).astype(dtype)
    hands_meanl = (0.12 * rng.normal(size=45)).astype(dtype)
    hands_meanr = (0.12 * rng.normal(size=45)).astype(dtype)

    kintree = np.stack([np.where(parents < 0, 2 ** 32 - 1, parents).astype(np.uint32),
                        np.arange(NUM_JOINTS, dtype=np.uint32)])
    return {
        'v_template': verts.astype(dtype),
        'f': faces,
        'shapedirs': shapedirs,
        'posedirs': posedirs,
        'J_regressor': Jreg.astype(dtype),
        'weights': W.astype(dtype),
        'kintree_table': kintree,
        'hands_componentsl': hands_componentsl,
        'hands_componentsr': hands_componentsr,
        'hands_meanl': hands_meanl,
        'hands_meanr': hands_meanr,
        'lmk_faces_idx': lmk_faces_idx.astype(np.int64),
        'lmk_bary_coords': lmk_bary.astype(dtype),
        'dynamic_lmk_faces_idx': dyn_faces.astype(np.int64),
        'dynamic_lmk_bary_coords': dyn_bary.astype(dtype),
    }


_CACHE = {}


def cached_smplx_like(seed=0, pose_corrective_scale=1.0):
    """Process-local cache (generation takes a few seconds).

    ``pose_corrective_scale`` rescales ``posedirs``.  The pose correctives of this module are
    white noise per vertex (up to ~2 cm at a strongly bent joint, the size of a triangle): fine
    for the keypoint terms, but it turns the surface around bent joints into a soup of crossing
    triangles, tens of thousands of pairs, which no real body mesh shows.  The interpenetration
    workload (BASELINE config 4) therefore uses 0.1 (millimetre-level correctives)."""
    key = (seed, float(pose_corrective_scale))
    if key not in _CACHE:
        if (seed, 1.0) not in _CACHE:
            _CACHE[(seed, 1.0)] = make_smplx_like(seed=seed, full_shape_space=False)
        if key not in _CACHE:
            m = dict(_CACHE[(seed, 1.0)])
            m['posedirs'] = (m['posedirs'] * np.float32(pose_corrective_scale)).astype(
                m['posedirs'].dtype)
            _CACHE[key] = m
    return _CACHE[key]


def without_degenerate_faces(model_data):
    """Copy of ``model_data`` whose faces with a repeated corner (the tube-man's cap fans hold 98
    of them) are replaced by a copy of the preceding proper face.  A triangle without area has no
    circumscribed cone: the reference's penalty (and ours) differentiates it to NaN.  With
    FilterFaces those faces never pair up on the test poses; without it they do, which no licensed
    body mesh can show -- the unfiltered interpenetration tests use this copy."""
    f = np.array(model_data['f'])
    bad = (f[:, 0] == f[:, 1]) | (f[:, 1] == f[:, 2]) | (f[:, 0] == f[:, 2])
    last = None
    for i in range(f.shape[0]):
        if bad[i]:
            if last is None:
                last = int(np.nonzero(~bad)[0][0])
            f[i] = f[last]
        else:
            last = i
    m = dict(model_data)
    m['f'] = f
    return m


def parts_segm_like(model_data):
    """Face segmentation with the keys of ``smplx_parts_segm.pkl`` (reference
    fit_single_frame.py:317-328) for a model of this module: the real file describes the
    licensed mesh topology, not the tube-man's.  ``segm[f]`` = bone owning the face (majority of
    its corners' dominant skinning joint), ``parents[f]`` = kinematic parent of that bone."""
    W = np.asarray(model_data['weights'])
    faces = np.asarray(model_data['f']).astype(np.int64)
    parents = np.asarray(model_data['kintree_table'])[0].astype(np.int64).copy()
    parents[0] = -1
    dom = W.argmax(1)[faces]
    segm = np.where(dom[:, 1] == dom[:, 2], dom[:, 1], dom[:, 0]).astype(np.int64)
    par = parents[segm]
    # the tube caps contain a few faces with a repeated corner (zero area, no circumcircle):
    # they get a part of their own that sibling_part_pairs() excludes from every pairing
    degen = ((faces[:, 0] == faces[:, 1]) | (faces[:, 1] == faces[:, 2]) |
             (faces[:, 0] == faces[:, 2]))
    segm[degen] = NUM_JOINTS
    par[degen] = -1
    return {'segm': segm, 'parents': par}


def sibling_part_pairs(model_data):
    """``ign_part_pairs`` entries for every pair of bones with the same parent: the tubes of
    sibling bones overlap at the shared joint by construction (the licensed mesh is one
    watertight surface and has no such overlaps)."""
    parents = np.asarray(model_data['kintree_table'])[0].astype(np.int64).copy()
    parents[0] = -1
    n = parents.shape[0]
    pairs = ['{},{}'.format(a, b) for a in range(n) for b in range(a + 1, n)
             if parents[a] == parents[b]]
    return pairs + ['{},{}'.format(a, n) for a in range(n)]       # zero-area faces: never paired


# Part pairs whose tubes already cross in the rest pose of make_smplx_like (found with the
# restated search of oracle/isect_port.py on the template; sibling pairs excluded).  The reference
# lists such pairs of ITS mesh in ign_part_pairs ("Upper arms and Spine 2", "Neck and jaw" in the
# yaml files); for the tube-man the list is longer.
REST_POSE_TOUCHING_PART_PAIRS = [
    '1,5', '4,10', '5,11', '9,16', '9,17', '12,16', '12,17', '12,22', '13,17', '25,29', '28,35',
    '32,34', '40,44', '43,50', '47,49']


def make_vposer_like(seed=2, latent=32, hidden=512, dtype=np.float32):
    """Weights with the VPoser v1 layer shapes, seeded.

    decoder: latent -> hidden -> hidden -> 21*6 ; encoder: BN(63) -> FC -> BN -> FC -> (mu, logvar)
    """
    rng = np.random.default_rng(seed)

    def lin(i, o, scale=1.0):
        w = rng.normal(size=(o, i)) * (scale / np.sqrt(i))
        b = rng.normal(size=(o,)) * 0.01
        return w.astype(dtype), b.astype(dtype)
    out = {}
    out['dec_fc1_w'], out['dec_fc1_b'] = lin(latent, hidden)
    out['dec_fc2_w'], out['dec_fc2_b'] = lin(hidden, hidden)
    w, b = lin(hidden, 126, scale=0.3)
    # bias the 6-D outputs towards the identity rotation so decoded poses are mild
    b = b.reshape(21, 3, 2)
    b[:, 0, 0] += 1.0
    b[:, 1, 1] += 1.0
    out['dec_out_w'], out['dec_out_b'] = w, b.reshape(-1).astype(dtype)
    out['enc_bn1_mean'] = np.zeros(63, dtype)
    out['enc_bn1_var'] = np.full(63, 0.09, dtype)
    out['enc_bn1_w'] = np.ones(63, dtype)
    out['enc_bn1_b'] = np.zeros(63, dtype)
    out['enc_fc1_w'], out['enc_fc1_b'] = lin(63, hidden)
    out['enc_bn2_mean'] = np.zeros(hidden, dtype)
    out['enc_bn2_var'] = np.ones(hidden, dtype)
    out['enc_bn2_w'] = np.ones(hidden, dtype)
    out['enc_bn2_b'] = np.zeros(hidden, dtype)
    out['enc_fc2_w'], out['enc_fc2_b'] = lin(hidden, hidden)
    out['enc_mu_w'], out['enc_mu_b'] = lin(hidden, latent)
    out['enc_logvar_w'], out['enc_logvar_b'] = lin(hidden, latent, scale=0.1)
    return out


def make_gmm_like(seed=1, num_gaussians=8, dim=63):
    """A dict with the ``gmm_08.pkl`` keys: means [M,d], covars [M,d,d] (SPD), weights [M]."""
    rng = np.random.default_rng(seed)
    means = 0.15 * rng.normal(size=(num_gaussians, dim))
    covars = np.zeros((num_gaussians, dim, dim))
    for m in range(num_gaussians):
        a = rng.normal(size=(dim, dim)) * 0.05
        covars[m] = a @ a.T + np.eye(dim) * (0.02 + 0.01 * m)
    w = rng.uniform(0.5, 1.5, size=num_gaussians)
    return {'means': means, 'covars': covars, 'weights': w / w.sum()}
