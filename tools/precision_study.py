#!/usr/bin/env python
"""Which rounding sets the output error of the PDA transformer block in a 16-bit-operand mode?  (CPU, float64 reference.)

The block: qkv = in_proj(y); ctx = attention(q, k, v); z = LN2(y + out_proj(ctx)); out = max_s(z + lin2(relu(lin1(z)))).
Each tensor is rounded to fp16 ALONE ("only x") and then in groups, against the float64 block with nn.MultiheadAttention /
nn.Linear default initialisation.  Result (d_model 256, nsample 32; printed below): rounding the RESIDUAL streams y and z
costs 3e-4 each, every GEMM operand (activations AND weights) together 8e-5.  Hence the fp16 single-pass mode of
csrc/tc_gemm.cu (NPASS = 4): one fp16 x fp16 MMA per k-step for every product, residual streams kept as (hi, lo) fp16 plane
pairs (22 bits) — the same output error class as the split-bf16 mode at a third of the MMAs and half the HBM bytes.

    python tools/precision_study.py
"""
import torch, math
def q16(x): return x.half().double()
def split2(x):
    hi = x.half().double(); lo=(x-hi).half().double(); return hi+lo
ident = lambda x: x
def run(E, ns, G, Q):
    heads=4; hd=E//heads
    mha = torch.nn.MultiheadAttention(E, heads).double()
    with torch.no_grad():
        mha.in_proj_bias.normal_(0, 0.02); mha.out_proj.bias.normal_(0,0.02)
    lin1 = torch.nn.Linear(E, E//2).double(); lin2 = torch.nn.Linear(E//2, E).double()
    ln2 = torch.nn.LayerNorm(E).double()
    T = G*ns
    tok = torch.randn(T, E).double()*torch.rand(1,E).double()*3
    y = torch.nn.functional.layer_norm(tok, (E,))
    g = lambda k: Q.get(k, ident)
    W = g("w")
    y_res = g("y_res")(y); y_a = g("y_a")(y)
    qkv = y_a @ W(mha.in_proj_weight.detach()).t() + mha.in_proj_bias.detach()
    q,k,v = qkv.split(E, dim=1)
    def hs(t): return t.view(G, ns, heads, hd).permute(0,2,1,3)
    q,k,v = g("qk")(q),g("qk")(k),g("v")(v)
    s = hs(q) @ hs(k).transpose(-1,-2) / math.sqrt(hd)
    p = g("p")(torch.softmax(s, -1))
    ctx = g("ctx")((p @ hs(v)).permute(0,2,1,3).reshape(T, E))
    z = ln2(y_res + ctx @ W(mha.out_proj.weight.detach()).t() + mha.out_proj.bias.detach()).detach()
    z_res = g("z_res")(z); z_a = g("z_a")(z)
    h = g("h")(torch.relu(z_a @ W(lin1.weight.detach()).t() + lin1.bias.detach()))
    o = z_res + h @ W(lin2.weight.detach()).t() + lin2.bias.detach()
    return o.view(G, ns, E).max(1)[0]
E, ns = 256, 32
torch.manual_seed(1); ref = run(E, ns, 64, {})
allk = ["w","y_res","y_a","qk","v","p","ctx","z_res","z_a","h"]
for k in allk:
    torch.manual_seed(1); o = run(E, ns, 64, {k: q16})
    e=(o-ref).abs(); print("only", k, "max/scale %.2e rms %.2e"%((e.max()/ref.abs().max()).item(), (e.pow(2).mean().sqrt()/ref.pow(2).mean().sqrt()).item()))
for name, ks in (("all", allk), ("all but residuals", [k for k in allk if k not in ("y_res","z_res")]),
                 ("all but z_res", [k for k in allk if k!="z_res"]), ("acts only no w/res",[k for k in allk if k not in ("w","y_res","z_res")])):
    torch.manual_seed(1); o = run(E, ns, 64, {k: q16 for k in ks})
    e=(o-ref).abs(); print(name, "max/scale %.2e rms %.2e"%((e.max()/ref.abs().max()).item(), (e.pow(2).mean().sqrt()/ref.pow(2).mean().sqrt()).item()))
# activations hi/lo split (two planes) but weights single fp16
torch.manual_seed(1); o = run(E, ns, 64, {"w": q16})
