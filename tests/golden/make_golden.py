"""Generate golden vectors for the MSDeformAttn path FROM THE REFERENCE ITSELF.

Run in the authoring container only (needs /root/reference, which does not exist on the GPU
box):  ``python tests/golden/make_golden.py``.  It imports the reference's own
``ms_deform_attn_core_pytorch`` (OPS/functions/ms_deform_attn_func.py:52-72) -- with an empty
stub for the un-built ``MultiScaleDeformableAttention`` extension that the file imports at
:21-29 -- runs it (and its autograd backward) on seeded inputs and writes the small ``.npz``
fixtures that ``tests/test_oracle.py`` / ``tests/test_gpu_parity.py`` check against.

Cases
  kat_opstest_D{2,30,32,64,71}  the OPS/test.py recipe (:24-39: N,M=1,2; Lq,L,P=2,2,2;
                                shapes [(6,4),(3,2)]; torch.manual_seed(3); value=rand*0.01,
                                loc=rand, attn=(rand+1e-5) normalised), fp64, for the channel
                                counts of OPS/test.py:88 that stay small.
  ragged_O_f64 / ragged_O_f32   non-square ragged pyramid, Lq != S, P=3, out-of-range locations
  cfg1_R_f32                    BASELINE config 1 (N=2, CVPPP 512 pyramid, d256, 8 heads,
                                4 points, realistic locations, seed 0): inputs are regenerated
                                from pctrans_b200.synth (sha256 of the bytes is stored), outputs
                                and grads are stored sub-sampled.
"""
import hashlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_FUNC = ("/root/reference/connectomics/model/maskformer_block/pixel_decoder/ops/"
            "functions/ms_deform_attn_func.py")


def load_reference_core():
    sys.modules.setdefault("MultiScaleDeformableAttention",
                           types.ModuleType("MultiScaleDeformableAttention"))
    spec = importlib.util.spec_from_file_location("_ref_ms_deform_attn_func", REF_FUNC)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ms_deform_attn_core_pytorch


def run_reference(core, value, shapes, loc, attn, grad_out):
    value = value.clone().requires_grad_(True)
    loc = loc.clone().requires_grad_(True)
    attn = attn.clone().requires_grad_(True)
    out = core(value, shapes, loc, attn)
    out.backward(grad_out)
    return out.detach(), value.grad, loc.grad, attn.grad


def sha(*tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.contiguous().numpy().tobytes())
    return h.hexdigest()


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: (v.numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrays.items()})
    print(f"wrote {path} ({os.path.getsize(path)} B)")


def main():
    from pctrans_b200 import synth

    core = load_reference_core()
    torch.set_num_threads(os.cpu_count() or 1)

    # --- OPS/test.py recipe ---------------------------------------------------------------
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    N, M, Lq, L, P = 1, 2, 2, 2, 2
    for D in (2, 30, 32, 64, 71):
        torch.manual_seed(3)
        value = (torch.rand(N, S, M, D) * 0.01).double()
        loc = torch.rand(N, Lq, M, L, P, 2).double()
        attn = torch.rand(N, Lq, M, L, P) + 1e-5
        attn = (attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)).double()
        grad_out = torch.rand(N, Lq, M * D).double()
        out, gv, gl, ga = run_reference(core, value, shapes, loc, attn, grad_out)
        save(f"kat_opstest_D{D}", value=value, shapes=shapes, lsi=lsi, loc=loc, attn=attn,
             grad_out=grad_out, out=out, grad_value=gv, grad_loc=gl, grad_attn=ga)

    # --- ragged, out-of-range -------------------------------------------------------------
    rag = ((5, 7), (3, 4), (2, 2))
    for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        value, shapes, lsi, loc, attn = synth.make_inputs(2, rag, M=3, D=8, P=3, regime="O",
                                                          seed=11, Lq=11)
        g = torch.Generator().manual_seed(12)
        grad_out = torch.randn(2, 11, 3 * 8, generator=g)
        value, loc, attn, grad_out = (t.to(dt) for t in (value, loc, attn, grad_out))
        out, gv, gl, ga = run_reference(core, value, shapes, loc, attn, grad_out)
        save(f"ragged_O_{tag}", value=value, shapes=shapes, lsi=lsi, loc=loc, attn=attn,
             grad_out=grad_out, out=out, grad_value=gv, grad_loc=gl, grad_attn=ga)

    # --- BASELINE config 1, sub-sampled ---------------------------------------------------
    value, shapes, lsi, loc, attn = synth.make_inputs(2, synth.CVPPP_512, M=8, D=32, P=4,
                                                      regime="R", seed=0)
    g = torch.Generator().manual_seed(1)
    grad_out = torch.randn(2, value.shape[1], 256, generator=g)
    out, gv, gl, ga = run_reference(core, value, shapes, loc, attn, grad_out)
    stride = 257
    save("cfg1_R_f32", input_sha256=np.frombuffer(
        bytes.fromhex(sha(value, loc, attn, grad_out)), dtype=np.uint8),
         stride=np.int64(stride),
         out=out.reshape(-1)[::stride], grad_value=gv.reshape(-1)[::stride],
         grad_loc=gl.reshape(-1)[::stride], grad_attn=ga.reshape(-1)[::stride],
         out_absmax=out.abs().max(), grad_value_absmax=gv.abs().max(),
         grad_loc_absmax=gl.abs().max(), grad_attn_absmax=ga.abs().max())


if __name__ == "__main__":
    main()
