"""-m gpu: failure behaviour of the product path (SURVEY.md section 8(b) "Errors").

The reference's callers catch exceptions and fall back to the input face (inference_codeformer.py:209-211; the Gradio demo
catches RuntimeError only, web-demos/hugging_face/app.py:176), so a failure inside a kernel must become a RuntimeError and
must leave the CUDA context usable: no trap, no sticky error, no out-of-bounds access on NaN input, no silent fp16 inf."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import codeformer_b200 as cb
from codeformer_b200 import _lib
from codeformer_b200 import spec as S
from tests import gpu_util as G
from tests.util import faces_input, golden, maxabs

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


def _conv_case():
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 128, 64, 64, generator=g)
    w = torch.randn(128, 128, 3, 3, generator=g) / (9 * 128) ** 0.5
    b = torch.randn(128, generator=g)
    return x, w, b, F.conv2d(x, w, b, padding=1)


def test_injected_pipeline_timeout_is_recoverable():
    """A TMA load that never arrives (injected) makes the MMA warp's barrier wait time out.  That must surface as a
    RuntimeError through the status word, and the very next launches -- the same conv, then a whole forward -- must work."""
    lib = _lib.load()
    x, w, b, ref = _conv_case()
    ok = G.conv2d(x, w, b, engine=2)                          # healthy run first (also binds the status word on this device)
    assert maxabs(ok.cpu(), ref) < 2e-5 * float(ref.abs().max()) + 1e-5
    cb.check_async_status()                                   # nothing pending
    _lib.check(lib.cfb_debug_set_wait_limit(40_000_000), 'set_wait_limit')     # ~25 ms instead of ~2 s
    try:
        _lib.check(lib.cfb_debug_inject_fault(1), 'inject_fault')
        G.conv2d(x, w, b, engine=2)                           # returns (the kernel aborts its pipeline instead of hanging)
        with pytest.raises(RuntimeError, match='time-out'):
            cb.check_async_status()
        cb.check_async_status()                               # reported once, then cleared
    finally:
        _lib.check(lib.cfb_debug_set_wait_limit(4_000_000_000), 'set_wait_limit')
    again = G.conv2d(x, w, b, engine=2)
    assert torch.equal(again, ok), 'the context must be fully usable after a reported time-out'
    net = cb.CodeFormer().cuda().eval()
    net.load_state_dict(S.random_state_dict(S.codeformer_spec(), 1))
    g = golden('codeformer_main.npz')
    out, logits, _ = net(faces_input(slice(0, 1)).cuda(), w=0.5, adain=True)
    torch.cuda.synchronize()
    cb.check_async_status()
    assert np.array_equal(logits.argmax(2).cpu().numpy(), g['top_idx']) and maxabs(out.cpu(), g['out']) < 1e-3


def test_failure_inside_restore_faces_falls_back_to_the_input_face(monkeypatch):
    """The batched caller front-end mirrors the reference's per-face fallback: a reported kernel failure returns the INPUT
    faces of that chunk (on_error='input') or raises (on_error='raise'); the following call is healthy."""
    monkeypatch.setenv('CFB_CUDA_GRAPH', '0')      # the fault hook acts on a host-side launch, not on a graph replay
    lib = _lib.load()
    net = cb.CodeFormer().cuda().eval()
    net.load_state_dict(S.random_state_dict(S.codeformer_spec(), 1))
    faces = golden('faces.npz')['faces'][:2][..., ::-1].copy()               # uint8 BGR
    good = net.restore_faces(faces, w=0.5, adain=True, on_error='raise')
    _lib.check(lib.cfb_debug_set_wait_limit(40_000_000), 'set_wait_limit')
    try:
        _lib.check(lib.cfb_debug_inject_fault(1), 'inject_fault')
        got = net.restore_faces(faces, w=0.5, adain=True, on_error='input')
        assert len(net.last_restore_errors) == 1 and 'time-out' in net.last_restore_errors[0][1]
        assert all(np.array_equal(a, b) for a, b in zip(got, faces)), 'fallback = the input faces'
        _lib.check(lib.cfb_debug_inject_fault(1), 'inject_fault')
        with pytest.raises(RuntimeError):
            net.restore_faces(faces, w=0.5, adain=True, on_error='raise')
    finally:
        _lib.check(lib.cfb_debug_set_wait_limit(4_000_000_000), 'set_wait_limit')
    after = net.restore_faces(faces, w=0.5, adain=True, on_error='raise')
    assert all(np.array_equal(a, b) for a, b in zip(after, good))


def test_nan_input_is_not_a_fault():
    """torch.argmax / topk return a valid index on NaN rows; so must the code lookup and the VQ argmin (an out-of-bounds
    gather would be a sticky context error).  NaN in -> NaN out, and the next call is healthy."""
    net = cb.CodeFormer().cuda().eval()
    net.load_state_dict(S.random_state_dict(S.codeformer_spec(), 1))
    x = faces_input(slice(0, 1)).cuda()
    ref = net(x, w=0.5, adain=True)[0].clone()
    bad = x.clone()
    bad[0, :, 100:110, 100:110] = float('nan')
    out, logits, _ = net(bad, w=0.5, adain=True)
    torch.cuda.synchronize()
    assert out.shape == ref.shape
    assert torch.equal(net(x, w=0.5, adain=True)[0], ref)
    vq = cb.VectorQuantizer(1024, 256, 0.25).cuda()
    z = torch.full((1, 256, 16, 16), float('nan'), device='cuda')
    zq, _, st = vq(z)
    torch.cuda.synchronize()
    idx = st['min_encoding_indices']
    assert int(idx.min()) >= 0 and int(idx.max()) < 1024
    z2 = torch.full((1, 256, 8, 8), float('inf'), device='cuda')                # off the tensor path (CUDA-core variant)
    _, _, st2 = vq(z2)
    torch.cuda.synchronize()
    assert int(st2['min_encoding_indices'].min()) >= 0 and int(st2['min_encoding_indices'].max()) < 1024


@pytest.mark.parametrize('gain', [1e3, 5e3, 5e4])
def test_fp16_operand_range_guard(gain):
    """Operands of the tensor-core path are fp16 pairs: |activation| must stay below 65504 where a RAW tensor is split
    (DESIGN.md section 4 "Range").  With the first conv scaled by `gain` the encoder's residual stream reaches 2e3, 1e4 and 9e4.
    Contract: either the result still meets the 1e-3 bar against the oracle, or a RuntimeError names the overflow -- never a
    silent inf/NaN."""
    from oracle import codeformer_oracle as O
    sd = S.random_state_dict(S.codeformer_spec(), 1)
    sd['encoder.blocks.0.weight'] = sd['encoder.blocks.0.weight'] * gain
    sd['encoder.blocks.0.bias'] = sd['encoder.blocks.0.bias'] * gain
    x = faces_input(slice(0, 1))
    col = {}
    ro, rl, _ = O.codeformer_forward(sd, x, w=0.5, adain_on=True, collect=col)
    stream = float(col['enc.2'].abs().max())
    net = cb.CodeFormer().cuda().eval()
    net.load_state_dict(sd)
    raised = None
    try:
        out, logits, _ = net(x.cuda(), w=0.5, adain=True)
        torch.cuda.synchronize()
        cb.check_async_status()
    except RuntimeError as e:
        raised = str(e)
    print(f'gain {gain:g}: residual stream max {stream:.3e}; ' + (f'raised: {raised[:120]}' if raised else
          f'out err {maxabs(out.cpu(), ro):.3e} logits err {maxabs(logits.cpu(), rl):.3e}'))
    if raised is None:
        assert bool(torch.isfinite(out).all()), 'silent inf/NaN'
        assert torch.equal(logits.argmax(2).cpu(), rl.argmax(2))
        assert maxabs(out.cpu(), ro) < 1e-3
    else:
        assert 'fp16' in raised and stream > 6e4, 'an overflow may only be reported when a raw operand really left the fp16 range'
    healthy = cb.CodeFormer().cuda().eval()
    healthy.load_state_dict(S.random_state_dict(S.codeformer_spec(), 1))
    g = golden('codeformer_main.npz')
    o2 = healthy(x.cuda(), w=0.5, adain=True)[0]
    torch.cuda.synchronize()
    cb.check_async_status()
    assert maxabs(o2.cpu(), g['out']) < 1e-3


def test_net_follows_its_device_and_second_device():
    """One process may drive several GPUs (per-device function attributes, per-device slab): moving a net re-prepares it on
    the new device, and a forward issued for another device than the net's raises instead of touching foreign memory."""
    net = cb.CodeFormer().cuda().eval()
    net.load_state_dict(S.random_state_dict(S.codeformer_spec(), 1))
    x = faces_input(slice(0, 1))
    ref = net(x.cuda(), w=0.5, adain=True)[0].cpu()
    if torch.cuda.device_count() < 2:
        pytest.skip('single-GPU box: only the device bookkeeping of device 0 ran')
    net.to('cuda:1')
    o1 = net(x.to('cuda:1'), w=0.5, adain=True)[0]
    assert torch.equal(o1.cpu(), ref)
    net.to('cuda:0')
    assert torch.equal(net(x.cuda(), w=0.5, adain=True)[0].cpu(), ref)
