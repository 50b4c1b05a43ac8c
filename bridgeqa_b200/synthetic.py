"""Deterministic ScanNet-shaped synthetic scenes (there are no datasets on the benchmark
machine).  Generated on the CPU with a seeded torch.Generator so the oracle, the
reference extension and the sm_100a kernels all see identical bits.

A scene is a room: x,y in [-4,4] m, z in [0,3] m; points sampled on the floor, the four
walls and ~20 axis-aligned boxes (furniture), plus N(0, 5 mm) noise, shuffled.  To
exercise the corner cases the reference kernels have, every scene also contains
  * 1 % exact duplicates (the ScanNet loader samples with replacement when a scan has
    fewer than num_points vertices, /root/reference/utils/pc_utils.py:25-34) -> real ties
    in FPS and ball query;
  * 8 points with |p|^2 <= 1e-3, which FPS must skip (sampling_gpu.cu:100-101).
Channel layout follows /root/reference/lib/dataset.py:380-412:
  xyz(3) | rgb(3) | normal(3) | height(1)                      -> C = 7   (config 2)
  xyz(3) | rgb(3) | height(1) | multiview(128)                 -> C = 132 (config 3)
"""
import torch

MEAN_COLOR_RGB = (109.8, 97.2, 83.8)


def _surface_points(gen, n):
    """n points on floor / walls / boxes of one room, (n,3) float32, plus unit normals."""
    kind = torch.rand(n, generator=gen)
    u = torch.rand(n, 3, generator=gen)
    pts = torch.empty(n, 3)
    nrm = torch.zeros(n, 3)
    # floor 35 %
    floor = kind < 0.35
    pts[floor] = torch.stack((u[floor, 0] * 8 - 4, u[floor, 1] * 8 - 4, torch.zeros(int(floor.sum()))), 1)
    nrm[floor, 2] = 1.0
    # walls 25 %
    wall = (kind >= 0.35) & (kind < 0.60)
    nw = int(wall.sum())
    which = torch.randint(0, 4, (nw,), generator=gen)
    t = u[wall, 0] * 8 - 4
    zz = u[wall, 1] * 3
    wx = torch.where(which == 0, torch.full((nw,), -4.0), torch.where(which == 1, torch.full((nw,), 4.0), t))
    wy = torch.where(which == 2, torch.full((nw,), -4.0), torch.where(which == 3, torch.full((nw,), 4.0), t))
    pts[wall] = torch.stack((wx, wy, zz), 1)
    wn = torch.zeros(nw, 3)
    wn[which == 0, 0] = 1.0
    wn[which == 1, 0] = -1.0
    wn[which == 2, 1] = 1.0
    wn[which == 3, 1] = -1.0
    nrm[wall] = wn
    # furniture 40 %: ~20 boxes, points on their faces
    box = kind >= 0.60
    nb = int(box.sum())
    nbox = 20
    centers = torch.stack((torch.rand(nbox, generator=gen) * 7 - 3.5,
                           torch.rand(nbox, generator=gen) * 7 - 3.5,
                           torch.zeros(nbox)), 1)
    sizes = torch.rand(nbox, 3, generator=gen) * torch.tensor([1.5, 1.5, 1.2]) + 0.3
    centers[:, 2] = sizes[:, 2] / 2
    bi = torch.randint(0, nbox, (nb,), generator=gen)
    face = torch.randint(0, 5, (nb,), generator=gen)           # +-x, +-y, top
    local = (u[box] - 0.5) * sizes[bi]
    axis = torch.tensor([0, 0, 1, 1, 2])[face]
    sign = torch.tensor([1.0, -1.0, 1.0, -1.0, 1.0])[face]
    rows = torch.arange(nb)
    local[rows, axis] = sign * sizes[bi, axis] / 2
    pts[box] = centers[bi] + local
    bn = torch.zeros(nb, 3)
    bn[rows, axis] = sign
    nrm[box] = bn
    return pts, nrm


def make_scene(scene_index, num_points=40000, num_features=7, seed_base=1000):
    """One scene, (num_points, 3 + num_features) float32.  num_features in {0, 1, 3, 4, 6, 7, 132}."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed_base + int(scene_index))
    pts, nrm = _surface_points(gen, num_points)
    pts = pts + torch.randn(num_points, 3, generator=gen) * 0.005
    pts[:, 2].clamp_(min=0.0)
    perm = torch.randperm(num_points, generator=gen)
    pts, nrm = pts[perm], nrm[perm]
    ndup = num_points // 100
    if ndup > 0:
        src = torch.randint(0, num_points, (ndup,), generator=gen)
        dst = torch.randint(0, num_points, (ndup,), generator=gen)
        pts[dst], nrm[dst] = pts[src], nrm[src]
    ntiny = min(8, num_points // 8)
    if ntiny > 0:
        where = torch.randint(1, num_points, (ntiny,), generator=gen)   # never index 0
        pts[where] = (torch.rand(ntiny, 3, generator=gen) - 0.5) * 0.03  # |p|^2 <= 6.75e-4
    cols = [pts]
    rgb = (torch.rand(num_points, 3, generator=gen) * 255 - torch.tensor(MEAN_COLOR_RGB)) / 256.0
    floor_height = torch.quantile(pts[:, 2], 0.01)
    height = (pts[:, 2] - floor_height).unsqueeze(1)
    if num_features == 7:
        cols += [rgb, nrm, height]
    elif num_features == 132:
        cols += [rgb, height, torch.rand(num_points, 128, generator=gen)]
    elif num_features == 6:
        cols += [rgb, nrm]
    elif num_features == 4:
        cols += [rgb, height]
    elif num_features == 3:
        cols += [rgb]
    elif num_features == 1:
        cols += [height]
    elif num_features != 0:
        cols += [torch.rand(num_points, num_features, generator=gen)]
    return torch.cat(cols, 1).float().contiguous()


def make_batch(batch_size, num_points=40000, num_features=7, first_scene=0, seed_base=1000):
    """(B, N, 3 + C) float32 on the CPU; scene i uses seed seed_base + first_scene + i."""
    return torch.stack([make_scene(first_scene + i, num_points, num_features, seed_base)
                        for i in range(batch_size)], 0)


def randomize_bn_stats(module, seed=0):
    """Give every BatchNorm non-trivial running statistics (mean N(0,.1), var U[.5,1.5]) and
    affine parameters so that eval-mode folding is actually exercised."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    for m in module.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            c = m.num_features
            m.running_mean.copy_(torch.randn(c, generator=gen) * 0.1)
            m.running_var.copy_(torch.rand(c, generator=gen) + 0.5)
            with torch.no_grad():
                m.weight.copy_(torch.rand(c, generator=gen) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(c, generator=gen) * 0.1)
    return module


def fill_state_dict(module, seed=0):
    """Deterministic weights that do not depend on module construction order: every entry of
    module.state_dict() is drawn from its own generator seeded by (seed, key).  Conv / linear
    weights ~ N(0, 2/fan_in) (kaiming-normal, the reference's init, pytorch_utils.py:168),
    biases ~ N(0, .05), BN gamma U[.75,1.25], beta N(0,.1), running_mean N(0,.1),
    running_var U[.5,1.5].  Returns the module (loaded, strict)."""
    import zlib
    sd = module.state_dict()
    out = {}
    for key in sorted(sd):
        ref = sd[key]
        gen = torch.Generator(device="cpu")
        gen.manual_seed((int(seed) * 1000003 + zlib.crc32(key.encode())) % (2 ** 63))
        shape = tuple(ref.shape)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            val = torch.zeros(shape, dtype=ref.dtype)
        elif leaf == "running_mean":
            val = torch.randn(shape, generator=gen) * 0.1
        elif leaf == "running_var":
            val = torch.rand(shape, generator=gen) + 0.5
        elif leaf == "weight" and len(shape) == 1:
            val = torch.rand(shape, generator=gen) * 0.5 + 0.75
        elif leaf == "bias" and (key[:-5] + ".running_mean") in sd:
            val = torch.randn(shape, generator=gen) * 0.1
        elif leaf == "bias":
            val = torch.randn(shape, generator=gen) * 0.05
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            val = torch.randn(shape, generator=gen) * (2.0 / max(fan_in, 1)) ** 0.5
        out[key] = val.to(ref.dtype)
    module.load_state_dict(out, strict=True)
    return module
