"""Synthetic Kinect-V2-shaped stair clouds (BASELINE.json configs 3-5).

Back-projection restates K2G::updateCloud / prepareMake3D (/root/reference/modules/k2g/k2g.cpp:231-308, :506-520):
pixel (r, c) of the 424 x 512 depth image -> camera ray ((c - cx + 0.5)/fx, (r - cy + 0.5)/fy, 1) * depth.  The
reference reads the intrinsics from the device at run time; nominal Kinect V2 IR values are used here and stated.
Scene: ground + 4 steps (rise 0.11 m, run 0.28 m, width 1.0 m -- the dimensions measured on samples/0001) + a
landing; depth noise sigma(z) = 1.5 mm + 1.5 mm (z - 0.5)^2, quantised to 1 mm like the 16-bit depth PNG
(test/reader.cpp:128); 20-30 % invalid pixels in blobs + everything beyond 4.5 m; rgba = 0 marks invalid pixels and
their xyz keep finite garbage, as in the recorded fixtures.  Output is the UN-levelled camera cloud plus the
levelling rotation R (test/reader.cpp:228-271), i.e. what Reader hands to rotationCloudPoints in Kinect mode.
"""
import numpy as np

W, H = 512, 424
FX = FY = 365.46
CX, CY = 254.88, 205.40


def _rot_from_euler(pitch, roll, yaw):
    ar, ap, ay = np.float32(roll / 180.0 * np.pi), np.float32(pitch / 180.0 * np.pi), np.float32(yaw / 180.0 * np.pi)
    cr, sr, cp, sp, cy, sy = np.cos(ar), np.sin(ar), np.cos(ap), np.sin(ap), np.cos(ay), np.sin(ay)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]], dtype=np.float64)
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]], dtype=np.float64)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]], dtype=np.float64)
    m1 = np.array([[0, 0, 1], [-1, 0, 0], [0, -1, 0]], dtype=np.float64)
    m3 = np.array([[0, 0, -1], [1, 0, 0], [0, -1, 0]], dtype=np.float64)
    return m3 @ (Rx @ Ry @ Rz) @ m1


def make_frame(frame_id=0, rise=0.11, run=0.28, width=1.0, n_steps=4, jitter=True, dropout=0.25):
    """Returns (records[N] {x,y,z,rgba}, R9 float32).  Seed: default_rng(20181013 + frame_id)."""
    rng = np.random.default_rng(20181013 + int(frame_id))
    cam_h = 0.95 + (rng.uniform(-0.05, 0.05) if jitter else 0.0)          # camera height above ground (m)
    pitch = 27.0 + (rng.uniform(-5, 5) if jitter else 0.0)
    roll = (rng.uniform(-3, 3) if jitter else 0.0)
    yaw_w = (rng.uniform(-5, 5) if jitter else 0.0)                      # heading of the staircase w.r.t. the camera
    y0 = 0.9 + (rng.uniform(-0.05, 0.05) if jitter else 0.0)             # distance to the first riser

    # pick the sign convention of the IMU pitch so that the optical axis tilts DOWN (+x is gravity)
    R = _rot_from_euler(pitch, roll, 0.0)
    if (R @ np.array([0, 0, 1.0]))[0] < 0:
        R = _rot_from_euler(-pitch, roll, 0.0)
    cols = (np.arange(W, dtype=np.float64) - CX + 0.5) / FX
    rows = (np.arange(H, dtype=np.float64) - CY + 0.5) / FY
    dirs = np.stack(np.broadcast_arrays(cols[None, :], rows[:, None], np.ones((H, W))), axis=-1)[:, ::-1, :]   # mirrored (K2G(true))
    dw = dirs.reshape(-1, 3) @ R.T                                         # world directions (x down, y fwd, z lateral)
    cy_, sy_ = np.cos(np.deg2rad(yaw_w)), np.sin(np.deg2rad(yaw_w))
    # rotate the staircase about the vertical axis: express rays in stair coordinates
    ds = np.stack([dw[:, 0], cy_ * dw[:, 1] + sy_ * dw[:, 2], -sy_ * dw[:, 1] + cy_ * dw[:, 2]], axis=1)
    t_best = np.full(ds.shape[0], np.inf)

    def hit_x(xp, ylo, yhi, zlo, zhi):          # horizontal rectangle at height x = xp
        with np.errstate(divide="ignore", invalid="ignore"):
            t = xp / ds[:, 0]
        y, z = t * ds[:, 1], t * ds[:, 2]
        ok = (t > 0) & (y >= ylo) & (y <= yhi) & (z >= zlo) & (z <= zhi)
        np.minimum(t_best, np.where(ok, t, np.inf), out=t_best)

    def hit_y(yp, xlo, xhi, zlo, zhi):          # vertical rectangle facing the camera at y = yp
        with np.errstate(divide="ignore", invalid="ignore"):
            t = yp / ds[:, 1]
        x, z = t * ds[:, 0], t * ds[:, 2]
        ok = (t > 0) & (x >= xlo) & (x <= xhi) & (z >= zlo) & (z <= zhi)
        np.minimum(t_best, np.where(ok, t, np.inf), out=t_best)

    hw = width / 2
    hit_x(cam_h, -10, 10, -10, 10)                                         # ground
    for i in range(1, n_steps + 1):
        ylo = y0 + (i - 1) * run
        yhi = y0 + i * run if i < n_steps else y0 + i * run + 1.0          # last tread = landing
        hit_x(cam_h - i * rise, ylo, yhi, -hw, hw)
        hit_y(ylo, cam_h - i * rise, cam_h - (i - 1) * rise, -hw, hw)
    hit_y(y0 + n_steps * run + 1.0, -3, cam_h - n_steps * rise, -hw, hw)   # back wall behind the landing

    depth = t_best.copy()                                                   # camera z (ray has z = 1)
    valid = np.isfinite(depth) & (depth < 4.5) & (depth > 0.3)
    sigma = 0.0015 + 0.0015 * (np.where(valid, depth, 1.0) - 0.5) ** 2
    depth = depth + rng.normal(0.0, 1.0, depth.shape) * sigma
    depth = np.round(depth * 1000.0) / 1000.0
    # blob dropout: threshold a smoothed noise field
    nz = rng.normal(size=(H // 8 + 2, W // 8 + 2))
    nz = np.kron(nz, np.ones((8, 8)))[:H, :W]
    cs = np.cumsum(np.pad(nz, ((5, 4), (0, 0))), axis=0)       # 9-tap box blur, rows then columns
    nz = cs[9:] - cs[:-9]
    cs = np.cumsum(np.pad(nz, ((0, 0), (5, 4))), axis=1)
    nz = cs[:, 9:] - cs[:, :-9]
    thr = np.quantile(nz, dropout)
    valid &= (nz.reshape(-1) > thr)

    d = np.where(valid, depth, 1.0 + 0.5 * rng.random(depth.shape))        # finite garbage under invalid pixels
    pc = dirs.reshape(-1, 3) * d[:, None]
    rec = np.zeros(W * H, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgba", "<u4")])
    rec["x"], rec["y"], rec["z"] = pc[:, 0], pc[:, 1], pc[:, 2]
    rgb = rng.integers(1, 256, size=(W * H, 3), dtype=np.uint32)
    rgba = (np.uint32(255) << 24) | (rgb[:, 0] << 16) | (rgb[:, 1] << 8) | rgb[:, 2]
    rec["rgba"] = np.where(valid, rgba, 0).astype(np.uint32)
    return rec, R.astype(np.float32).reshape(9)


INTRINSICS = (FX, FY, CX, CY)


def make_depth_frame(frame_id=0, **kw):
    """The same scene as make_frame(frame_id), as what the sensor side delivers before K2G::updateCloud: a 424 x 512 uint16
    depth image in millimetres (0 = no return; un-mirrored, i.e. before the horizontal flip of K2G(true)), a registered colour
    image (b, g, r, x; the front-end's validity mask needs non-zero channels) and the levelling rotation."""
    rec, R = make_frame(frame_id, **kw)
    valid = ((rec["rgba"] & 255) != 0) & (((rec["rgba"] >> 8) & 255) != 0) & (((rec["rgba"] >> 16) & 255) != 0)
    depth = np.where(valid, np.round(rec["z"].astype(np.float64) * 1000.0), 0).astype(np.uint16).reshape(H, W)[:, ::-1].copy()
    bgrx = np.zeros((H, W, 4), dtype=np.uint8)
    rgba = rec["rgba"].reshape(H, W)[:, ::-1]
    bgrx[..., 0], bgrx[..., 1], bgrx[..., 2] = rgba & 255, (rgba >> 8) & 255, (rgba >> 16) & 255
    return depth, bgrx, R


def make_batch(n, first_id=0, **kw):
    recs, Rs = [], []
    for i in range(n):
        r, R = make_frame(first_id + i, **kw)
        recs.append(r)
        Rs.append(R)
    return np.stack(recs), np.stack(Rs)


def make_dense_scene(n_points=2_000_000, seed=0, rise=0.11, run=0.28, width=1.0, n_steps=4, sigma=0.002, invalid_frac=0.01):
    """BASELINE.json config 4: an UNORGANISED dense stair scene, already levelled (+x down, +y forward, +z lateral).
    Points are sampled uniformly (by area) on ground, treads, risers, the landing and a back wall, with isotropic
    Gaussian noise `sigma`; `invalid_frac` of them get rgba = 0 (validity mask).  Returns records[n_points] {x,y,z,rgba};
    use it with a context of width = n_points, height = 1 and normal mode 1 (kNN normals, a3')."""
    rng = np.random.default_rng(424242 + int(seed))
    cam_h, y0, hw = 0.95, 0.9, width / 2
    rects = []          # (origin, edge u, edge v)
    rects.append((np.array([cam_h, 0.3, -1.0]), np.array([0, y0 - 0.3, 0]), np.array([0, 0, 2.0])))                 # ground
    for i in range(1, n_steps + 1):
        ylo = y0 + (i - 1) * run
        depth = run if i < n_steps else run + 1.0                                                                      # last tread = landing
        rects.append((np.array([cam_h - i * rise, ylo, -hw]), np.array([0, depth, 0]), np.array([0, 0, width])))      # tread
        rects.append((np.array([cam_h - i * rise, ylo, -hw]), np.array([rise, 0, 0]), np.array([0, 0, width])))       # riser
    yb = y0 + n_steps * run + 1.0
    rects.append((np.array([cam_h - n_steps * rise - 1.0, yb, -hw]), np.array([1.0, 0, 0]), np.array([0, 0, width])))  # back wall
    areas = np.array([np.linalg.norm(np.cross(u, v)) for _, u, v in rects])
    which = rng.choice(len(rects), size=n_points, p=areas / areas.sum())
    a, b = rng.random(n_points), rng.random(n_points)
    O = np.stack([r[0] for r in rects])[which]
    U = np.stack([r[1] for r in rects])[which]
    V = np.stack([r[2] for r in rects])[which]
    P = O + a[:, None] * U + b[:, None] * V + rng.normal(0.0, sigma, size=(n_points, 3))
    rec = np.zeros(n_points, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgba", "<u4")])
    rec["x"], rec["y"], rec["z"] = P[:, 0], P[:, 1], P[:, 2]
    rgb = rng.integers(1, 256, size=(n_points, 3), dtype=np.uint32)
    rgba = (np.uint32(255) << 24) | (rgb[:, 0] << 16) | (rgb[:, 1] << 8) | rgb[:, 2]
    rec["rgba"] = np.where(rng.random(n_points) < invalid_frac, 0, rgba).astype(np.uint32)
    return rec


DENSE_PARAMS = {"voxel_x": 0.005, "voxel_y": 0.005, "voxel_z": 0.005}     # config 4: fine leaf (the GUI minimum is 0.01)
