"""The fused ResBlock-layer kernel (csrc/layer_tc.cu: gate GEMM -> tanh*sigmoid -> res|skip 1x1, the gated tile kept in shared
memory) against the two-launch path it replaces (tc_gemm_kernel<GATE> + tc_gemm_kernel<RES_SKIP>, whose parity against the float64
oracle is pinned by test_gpu_mixed.py / test_gpu_baseline_shapes.py).  Both round o to 16 bits and accumulate in fp32 in the same
order, so the two must agree BIT FOR BIT -- forward (model.py:317-347) and reverse (model.py:350-396), bf16 and fp16 operands,
ragged tile counts (odd number of 128-row tiles, rows past the utterance end), deep blocks whose conditioning projection runs
ahead, and 1- / 3-layer WaveNets (modules.py:161-186) where only some layers qualify for the fused kernel."""
import pytest
import torch

from oracle import flowavenet_oracle as O

pytestmark = pytest.mark.gpu

CASES = [
    # (n_block, n_flow, n_layer, upsample_scales, B, frames, dtype)
    (5, 6, 2, (8, 12), 3, 11, "bfloat16"),      # hparams8000 depth, short; 15 row tiles at block 0: the last pair has a dummy tile
    (5, 6, 2, (8, 12), 3, 67, "bfloat16"),      # block 0: 3 x 26 row tiles (25.1 -> ragged), odd pair count
    (5, 6, 2, (8, 12), 1, 83, "float16"),       # fp16 operands (the accurate gate)
    (8, 6, 2, (16, 16), 1, 9, "bfloat16"),      # hparams depth: blocks 5..7 take the conditioning projection computed ahead
    (5, 6, 2, (8, 12), 3, 139, "bfloat16"),     # block 0: 80 row-tile pairs -> the 4-CTA cluster variant (multicast weights), last pair half dummy
    (5, 6, 2, (8, 12), 1, 393, "float16"),      # block 0: 74 pairs = 37 cluster units (odd)
    (8, 6, 2, (16, 16), 1, 5168, "bfloat16"),   # the C4 shape: block 5 (conditioning projection computed ahead) has 81 row-tile pairs, several per CTA pair
    (2, 2, 3, (2, 2), 2, 300, "bfloat16"),      # 3 layers: the middle one (residual AND running skip) stays on the two-launch path
    (2, 2, 1, (2, 2), 2, 200, "bfloat16"),      # 1 layer: skip only, nothing staged in
]


@pytest.mark.parametrize("n_block,n_flow,n_layer,scales,B,frames,dtype", CASES)
def test_fused_layer_bit_equal_to_two_launches(n_block, n_flow, n_layer, scales, B, frames, dtype):
    import tf_flowavenet_b200 as P
    mels = 80 if len(scales) == 2 and scales[0] >= 8 else 8
    hp = O.HP(n_block=n_block, n_flow=n_flow, n_layer=n_layer, num_mels=mels, upsample_scales=scales)
    params = O.synthetic_params(hp, 5)
    x, c = O.synthetic_inputs(hp, B, frames, 6, "x")
    z_in, _ = O.synthetic_inputs(hp, B, frames, 7, "z")
    net = P.FloWaveNet(P.HParams(n_block=n_block, n_flow=n_flow, n_layer=n_layer, num_mels=mels, upsample_scales=list(scales), dtype=dtype),
                       variables=P.VariableStore())
    net.load_variables({k: v.numpy() for k, v in params.items()})
    xd, cd, zd = x.cuda(), c.cuda(), z_in.cuda()
    out = {}
    for mode in (0, 1):
        net.set_layer_fusion(mode)
        for rep in range(3):   # eager, graph capture, graph replay
            lp, ld, z = net.forward(xd, cd, return_z=True)
            xr = net.reverse(zd, cd)
        torch.cuda.synchronize()
        out[mode] = (float(lp), float(ld), z.clone(), xr.clone(), net.last_launches())
    assert torch.isfinite(out[1][2]).all() and torch.isfinite(out[1][3]).all()
    dz = float((out[0][2] - out[1][2]).abs().max())
    dx = float((out[0][3] - out[1][3]).abs().max())
    print("fused layer %s B=%d frames=%d: |dz| %.3g |dx| %.3g, log_p %.7f vs %.7f, launches per reverse pass %d -> %d" %
          (dtype, B, frames, dz, dx, out[0][0], out[1][0], out[0][4], out[1][4]))
    assert torch.equal(out[0][2], out[1][2]), dz
    assert torch.equal(out[0][3], out[1][3]), dx
    # log_p / log-det are double-precision atomic sums of identical terms: equal up to summation order
    assert abs(out[0][0] - out[1][0]) <= 1e-6 * max(1.0, abs(out[0][0])) and abs(out[0][1] - out[1][1]) <= 1e-6 * max(1.0, abs(out[0][1]))
    assert out[1][4] < out[0][4]   # the fused path really ran (fewer launches)


def test_front_direct_tile_sizes_agree(monkeypatch):
    """front_direct_kernel (csrc/conv_simt.cu; modules.py:164-165 front conv straight from the flow variable) picks 256- / 128-row tiles
    for long inputs and 64-row tiles otherwise: both on the same input, bit for bit (ragged tiles, utterance borders inside a tile)."""
    import tf_flowavenet_b200 as P
    hp = O.HP(n_block=5, n_flow=6, n_layer=2, num_mels=80, upsample_scales=(8, 12))
    params = O.synthetic_params(hp, 15)
    x, c = O.synthetic_inputs(hp, 3, 41, 16, "x")     # block 0: 3 x 1968 rows = 7.7 tiles of 256, block 1: 3 x 984 rows
    z_in, _ = O.synthetic_inputs(hp, 3, 41, 17, "z")
    net = P.FloWaveNet(P.HParams(n_block=5, n_flow=6, n_layer=2, num_mels=80, upsample_scales=[8, 12], dtype="bfloat16"), variables=P.VariableStore())
    net.load_variables({k: v.numpy() for k, v in params.items()})
    xd, cd, zd = x.cuda(), c.cuda(), z_in.cuda()
    out = {}
    for mode in ("small", "big"):
        monkeypatch.setenv("FWN_FRONT_TILE", mode)
        net.set_layer_fusion(-1)   # drops the captured graphs: the next passes launch with the new setting
        for rep in range(2):
            lp, ld, z = net.forward(xd, cd, return_z=True)
            xr = net.reverse(zd, cd)
        torch.cuda.synchronize()
        out[mode] = (z.clone(), xr.clone())
    assert torch.isfinite(out["big"][0]).all()
    assert torch.equal(out["small"][0], out["big"][0]) and torch.equal(out["small"][1], out["big"][1])
