"""The two rewrites inside `heights_kernel`'s loop body (csrc/ti5_heights.cu) keep the sampled cell for every input:
restated in numpy float32 (one rounding per operation, like the library's -fmad=false build) next to the reference's
op chain — `quat_apply` of a yaw-only quaternion (isaacgym torch_utils) and `.long()` + `clip` (lr:1576-1581)."""
import numpy as np

F = np.float32


def _reference_chain(z, w, vx, vy, rx, ry, border, inv_scale):
    """quat_apply(q, v) = v + w * t + cross(q.xyz, t), t = 2 * cross(q.xyz, v), q.xyz = (0, 0, z), v = (vx, vy, 0)."""
    zero = F(0.0)
    tx = (zero * zero - z * vy) * F(2.0)
    ty = (z * vx - zero * zero) * F(2.0)
    tz = (zero * vy - zero * vx) * F(2.0)
    px = (vx + w * tx) + (zero * tz - z * ty)
    py = (vy + w * ty) + (z * tx - zero * tz)
    return ((px + rx) + border) * inv_scale, ((py + ry) + border) * inv_scale


def _kernel_chain(z, w, vx, vy, rx, ry, border, inv_scale):
    zero = F(0.0)
    tx = (zero - z * vy) * F(2.0)
    ty = (z * vx) * F(2.0)
    px = (vx + w * tx) + (zero - z * ty)
    py = (vy + w * ty) + (z * tx)
    return ((px + rx) + border) * inv_scale, ((py + ry) + border) * inv_scale


def _cell_reference(q, dim):
    """`.long()` truncates toward zero (NaN -> 0 on the GPU), then clip(0, dim - 2)."""
    with np.errstate(invalid="ignore"):
        t = np.where(np.isnan(q), 0.0, np.trunc(q.astype(np.float64)))
    return np.clip(t, 0, dim - 2).astype(np.int64)


def _cell_kernel(q, dim):
    """float clamp to [-1, dim - 1] (fmaxf / fminf drop a NaN operand), 32-bit truncation, integer clip."""
    c = np.fmin(np.fmax(q, F(-1.0)), F(dim - 1))
    return np.clip(np.trunc(c).astype(np.int32), 0, dim - 2).astype(np.int64)


def test_rotation_without_the_zero_products_rounds_the_same():
    rng = np.random.default_rng(7)
    n = 400_000
    yaw = rng.uniform(-np.pi, np.pi, n)
    z, w = np.sin(yaw / 2).astype(F), np.cos(yaw / 2).astype(F)
    z[:64], w[:64] = F(0.0), F(1.0)                         # no rotation: every product with z is an exact zero
    z[64:128], w[64:128] = F(1.0), F(0.0)
    vx = rng.choice(np.linspace(-0.8, 0.8, 17), n).astype(F)      # the scan grid of lr_cfg:29-30 (has exact zeros)
    vy = rng.choice(np.linspace(-0.5, 0.5, 11), n).astype(F)
    rx, ry = rng.uniform(-30, 30, n).astype(F), rng.uniform(-30, 30, n).astype(F)
    rx[:32] = F(-25.0)                                       # px + rx + border == 0 exactly for the unrotated centre point
    a = _reference_chain(z, w, vx, vy, rx, ry, F(25.0), F(1.0) / F(0.1))
    b = _kernel_chain(z, w, vx, vy, rx, ry, F(25.0), F(1.0) / F(0.1))
    for qa, qb in zip(a, b):
        assert np.array_equal(qa, qb)                        # equal as values: at most the sign of a zero differs
        nz = qa != 0
        assert np.array_equal(qa[nz].view(np.uint32), qb[nz].view(np.uint32))
        assert np.array_equal(_cell_reference(qa, 2100), _cell_kernel(qb, 2100))


def test_clamped_32_bit_truncation_picks_the_same_cell():
    rng = np.random.default_rng(11)
    dim = 2100
    q = np.concatenate([
        rng.uniform(-5, dim + 5, 500_000), rng.uniform(-1.5, 1.5, 100_000), rng.uniform(dim - 4, dim + 1, 100_000),
        np.arange(-3, dim + 3, dtype=np.float64), np.arange(-3, dim + 3) + 0.99999, np.arange(-3, dim + 3) - 0.00001,
        [np.nan, np.inf, -np.inf, 3e9, -3e9, 1e19, -1e19, 3.4e38, -3.4e38, 1e-45, -1e-45, 0.0, -0.0, -0.99999994, -1.0],
    ]).astype(F)
    assert np.array_equal(_cell_reference(q, dim), _cell_kernel(q, dim))
    for d in (2, 3, 17, (1 << 24) - 1):                      # smallest fields, and the largest the launch check admits
        assert np.array_equal(_cell_reference(q, d), _cell_kernel(q, d))
