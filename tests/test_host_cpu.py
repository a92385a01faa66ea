"""CPU-side checks (no GPU): the C-ABI library builds/loads and exports every symbol the public header
declares, the drop-in modules keep the reference's constructor surface and state_dict layout, and
nothing silently falls back to PyTorch on CPU."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mpgan_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mpg_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def libpath():
    from mpgan_b200 import build
    return build.build()


def test_header_symbols_exported(libpath):
    lib = ctypes.CDLL(libpath)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mpgan_b200.h but not exported"
    lib.mpg_version.restype = ctypes.c_int
    assert lib.mpg_version() >= 100
    lib.mpg_features.restype = ctypes.c_int
    assert lib.mpg_features() & 1, "tcgen05 edge forward must be compiled in"
    assert lib.mpg_features() & 2, "tcgen05 edge backward must be compiled in"


def test_binding_table_matches_header(libpath):
    from mpgan_b200 import _lib
    assert sorted(_lib.exported_symbols()) == _declared()


def test_library_has_no_torch_dependency(libpath):
    out = subprocess.run(["ldd", libpath], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libc10" not in out, out   # (bare "c10" also matches load addresses)


_SASS = {}


def _sass(libpath):
    """SASS of the library (one cuobjdump run per session)."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    if libpath not in _SASS:
        _SASS[libpath] = subprocess.run([cuobjdump, "-sass", libpath], capture_output=True, text=True).stdout
    return _SASS[libpath]


def test_sass_is_blackwell_native(libpath):
    """The edge kernels must be tcgen05 / TMEM / bulk-async code, not recompiled mma.sync."""
    sass = _sass(libpath)
    assert "UTCHMMA" in sass, "no tcgen05.mma in SASS"
    assert "LDTM" in sass, "no tcgen05.ld in SASS"
    assert "UBLKCP" in sass, "no bulk-async (TMA) copy in SASS"


def test_node_network_kernels_are_tcgen05(libpath):
    """The fused node-network kernels (fn_tc.cu) issue tcgen05.mma / tcgen05.ld / tcgen05.st themselves."""
    sass = _sass(libpath)
    cur, seen = None, {}
    for line in sass.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
        elif cur is not None:
            for op in ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "LDGSTS"):
                if op in line:
                    seen.setdefault(cur, set()).add(op)
    chain = [k for k in seen if "fn_tc_kernel" in k]
    dw = [k for k in seen if "fn_dw_kernel" in k]
    assert len(chain) == 2 and len(dw) == 1, (chain, dw)
    for k in chain:   # MMA, accumulator loads, in-place TMEM write-back of the next A operand, bulk weight copies
        assert {"UTCHMMA", "LDTM", "STTM", "UBLKCP"} <= seen[k], (k, seen[k])
    assert {"UTCHMMA", "LDTM", "LDGSTS"} <= seen[dw[0]], seen[dw[0]]   # cp.async staged MN-major operands


def test_workspace_queries_and_support_matrix():
    """Host-only entry points: no GPU needed."""
    from mpgan_b200 import _lib
    L = _lib.lib()
    fwd = L.mpg_edge_fwd_workspace_bytes(256, 30, 32, 96, 160, 192)
    full = L.mpg_edge_workspace_bytes(256, 30, 32, 96, 160, 192)
    assert 0 < fwd < full
    assert L.mpg_fn_supported(192, 32, 256, 256, 32, 0.5) == 1      # G layers, D layer 1
    assert L.mpg_fn_supported(192, 3, 256, 256, 32, 0.5) == 1       # D layer 0
    assert L.mpg_fn_supported(192, 32, 256, 256, 3, 0.0) == 1       # G's last layer
    assert L.mpg_fn_supported(192, 32, 256, 256, 32, 0.3) == 0      # other dropout rates: per-layer kernels
    assert L.mpg_fn_supported(192, 32, 200, 256, 32, 0.0) == 0      # other widths: per-layer kernels
    assert L.mpg_fn_supported(192, 100, 256, 256, 32, 0.0) == 0     # cat input wider than 256
    assert L.mpg_fn_workspace_bytes(192, 32, 256, 256, 32) >= (256 * 224 + 256 * 256 + 32 * 256) * 4


def test_state_dict_layout_matches_reference(golden):
    from mpgan_b200 import presets
    man = golden("mp_state_manifest.pt")
    G, D = presets.mp_generator(), presets.mp_discriminator()
    assert {k: tuple(v.shape) for k, v in G.state_dict().items()} == man["G"]
    assert {k: tuple(v.shape) for k, v in D.state_dict().items()} == man["D"]
    assert sum(p.numel() for p in G.parameters()) == 361123
    assert sum(p.numel() for p in D.parameters()) == 355617
    G.load_state_dict(golden("mp_g_weights.pt"), strict=True)
    D.load_state_dict(golden("mp_d_seed4_weights.pt"), strict=True)


def test_gapt_and_spectral_norm_layouts(golden):
    from mpgan_b200 import LinearNet, presets
    g = golden("gapt.pt")
    for name in ("sab", "isab"):
        GG = presets.gapt_generator(use_isab=name == "isab")
        GD = presets.gapt_discriminator(use_isab=name == "isab")
        GG.load_state_dict(g[name]["sdG"], strict=True)
        GD.load_state_dict(g[name]["sdD"], strict=True)
    sn = golden("spectral_norm.pt")
    net = LinearNet([24, 16], input_size=10, output_size=4, final_linear=True, spectral_norm=True)
    net.load_state_dict(sn["sd0"], strict=True)
    assert not net.net[0].module.weight_u.requires_grad and net.net[0].module.weight_bar.requires_grad
    assert "net.2.weight" in net.state_dict()  # final linear layer is not wrapped (reference :65-68)


def test_same_seed_same_init_as_reference_layout():
    """nn.Linear containers consume the RNG like the reference, so equal seeds give equal weights."""
    from mpgan_b200 import presets
    torch.manual_seed(4)
    a = presets.mp_discriminator().state_dict()
    torch.manual_seed(4)
    b = presets.mp_discriminator().state_dict()
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_no_cpu_fallback():
    from mpgan_b200 import presets
    G = presets.mp_generator()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G(torch.zeros(2, 30, 32), torch.ones(2, 1))
    D = presets.gapt_discriminator()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        D(torch.zeros(2, 30, 4))
    from mpgan_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.linear(torch.zeros(4, 8), torch.zeros(3, 8), torch.zeros(3), True, 0.2, 0.0)


def test_unsupported_options_fail_loudly():
    from mpgan_b200 import LinearNet, MPLayer, presets
    with pytest.raises(NotImplementedError):
        LinearNet([8], input_size=4, batch_norm=True)
    with pytest.raises(NotImplementedError):
        presets.mp_generator(mask_learn=True)
    with pytest.raises(ValueError):   # kNN feeds ONE distance column; the reference dies with a shape error here
        MPLayer(3, [96, 160, 192], [256, 256], 32, fully_connected=False, pos_diffs=True, delta_coords=True, delta_r=True)
    # supported since round 2: kNN message passing, conditioning columns, GAPT LayerNorm (reference parameter names)
    assert MPLayer(3, [96, 160, 192], [256, 256], 32, clabels=2, mask_fne_np=True).fe.net[0].weight.shape == (96, 9)
    assert presets.mp_discriminator(clabels=1).order_dependent and not presets.mp_discriminator().order_dependent
    l = MPLayer(3, [96, 160, 192], [256, 256], 32, fully_connected=False, num_knn=5, pos_diffs=True, all_ef=False)
    assert (l._ef_mode, l._nd, l.num_ef) == (1, 2, 1) and l.fe.net[0].weight.shape == (96, 7)
    g = presets.gapt_generator(layer_norm_gen=True)
    assert "sabs.0.mab.norm1.weight" in g.state_dict() and "sabs.3.mab.norm2.bias" in g.state_dict()
    with pytest.raises(RuntimeError):   # and still no CPU path
        l(torch.zeros(1, 4, 3))


def test_pair_feature_modes():
    from mpgan_b200 import MPLayer
    assert MPLayer(5, [16, 24, 32], [40], 6)._ef_mode == 0
    l = MPLayer(5, [16, 24, 32], [40], 6, pos_diffs=True, all_ef=True, delta_r=False)
    assert (l._ef_mode, l._nd, l.num_ef) == (1, 5, 1)
    l = MPLayer(5, [16, 24, 32], [40], 6, pos_diffs=True, all_ef=False, delta_r=True, delta_coords=True)
    assert (l._ef_mode, l._nd, l.num_ef) == (3, 2, 3)
    assert l.fe.net[0].weight.shape == (16, 13)


def test_flops_model_matches_survey():
    import bench
    fg, fd = bench.net_flops(30)
    assert abs(fg / 1e6 - 181.9) < 0.1 and abs(fd / 1e6 - 181.6) < 0.1
    assert abs(bench.step_flops(30) / 1e9 - 2.180) < 0.001
    assert abs(bench.step_flops(150) / 1e9 - 50.71) < 0.01


def test_sort_by_count_has_no_host_path():
    """train.sort_by_count orders the batch with the library's kernels for any batch size; CPU tensors raise (there is
    no torch detour inside the product path)."""
    from mpgan_b200 import train
    labels = torch.tensor([[0.5], [1.0], [0.5], [0.1], [1.0]])
    data = torch.arange(5, dtype=torch.float32).view(5, 1, 1).expand(5, 3, 4).contiguous()
    with pytest.raises(RuntimeError):
        train.sort_by_count(data, labels)


def test_rank_shard_covers_all_samples():
    from mpgan_b200 import train
    for n, w in ((1000, 8), (1001, 8), (5, 8), (1_000_000, 8), (7, 1)):
        spans = [train.rank_shard(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_compaction_map_reference_properties():
    """The layout rule of mpg_compact_map (numpy restatement, tests/compact_ref.py): every unmasked particle once, in
    order; with every jet at least 15 positions wide no 128-row tile touches more than 10 jets (the bound of the edge
    kernels' Q ring), for any mix of tiny, empty and full jets; and the executed-step count ops.edge_active_fraction
    derives from a map equals a direct count."""
    import numpy as np
    import torch
    from compact_ref import compact_map_ref, as_cmap
    from mpgan_b200 import ops
    rng = np.random.default_rng(0)
    for B, N, kind in [(64, 30, "uniform"), (200, 30, "tiny"), (33, 150, "uniform"), (50, 17, "holes"), (40, 30, "empty")]:
        if kind == "tiny":
            n = rng.integers(0, 4, B)
        elif kind == "empty":
            n = np.where(rng.random(B) < 0.5, 0, rng.integers(1, N + 1, B))
        else:
            n = rng.integers(1, N + 1, B)
        mask = (np.arange(N)[None, :] < n[:, None]).astype(np.float32)
        if kind == "holes":
            mask *= rng.random((B, N)) < 0.7
        ref = compact_map_ref(mask)
        rows = ref["rowmap"][ref["rowmap"] >= 0]
        assert np.array_equal(rows, np.nonzero(mask.reshape(-1))[0])
        assert ref["tiles"] <= ref["tmax"] and ref["tile_nj"].max() <= 10
        for t in range(ref["tiles"]):
            r = ref["rowmap"][t * 128:(t + 1) * 128]
            r = r[r >= 0]
            if r.size:
                assert ref["tile_j0"][t] == r[0] // N and ref["tile_j0"][t] + ref["tile_nj"][t] - 1 == r[-1] // N
            else:
                assert ref["tile_nj"][t] == 0
        # executed (tile, sender) steps: sender s of a tile is live if any of the tile's jets has it unmasked
        steps = 0
        for t in range(ref["tiles"]):
            j0, nj = int(ref["tile_j0"][t]), int(ref["tile_nj"][t])
            if nj:
                steps += int((mask[j0:j0 + nj] != 0).any(0).sum())
        frac = ops.edge_active_fraction(torch.from_numpy(mask), B, N, torch.from_numpy(as_cmap(ref)))
        assert abs(frac - steps / (((B * N + 127) // 128) * N)) < 1e-12
