"""Scratch: which tensor-core family carries the end-to-end error: convolutions or the attention kernels?"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import sma_b200 as S
import sma_oracle as O
from conftest import CFG
inv = json.load(open(os.path.join(ROOT, 'tests/golden/state_keys.json')))
P_g, P_me = O.synthetic_state_dict(inv['net_g'], 0), O.synthetic_state_dict(inv['motion_estimator'], 1)
g = S.build_network(CFG['network_g']); me = S.build_network(CFG['network_motion_estimator'])
g.load_state_dict(P_g); me.load_state_dict(P_me)
g, me = g.eval().cuda(), me.eval().cuda()
src, drv = O.synthetic_frames(64, seed=77)
sel = [drv[i] for i in (0, 17, 38, 63)]
with torch.no_grad():
    kp_s = O.kp_detector(P_me, src.unsqueeze(0)); kp_0 = O.kp_detector(P_me, drv[0].unsqueeze(0)); kp_d = O.kp_detector(P_me, torch.stack(sel))
    kp_n = O.normalize_kp(kp_s, kp_d, kp_0, True, True, True)
    dm = O.dense_motion(P_me, src.unsqueeze(0).expand(4, -1, -1, -1), kp_n, {k: v.expand(4, *v.shape[1:]) for k, v in kp_s.items()})
    ref = O.generator_forward(P_g, O.encode_source(P_g, src.unsqueeze(0)), dm, 1.0)
heat = S.ops.nchw_to_nhwc(dm['driving_kp_heatmap'].cuda().contiguous())
orig_mha, orig_conv, orig_gn, orig_ln = S.ops.mha, S.ops.conv2d, S.ops.groupnorm_stats, S.ops.layernorm
def run(tag):
    g.clear_source_cache()
    feats = g.encode_source(src.unsqueeze(0).cuda())
    r = g.generate(feats, dm['deformation'].cuda(), dm['occlusion_map'].cuda().view(4, 64, 64), heat, 1.0)
    e = (r['out'].permute(0, 3, 1, 2).cpu() - ref['out']).abs().amax(dim=(1, 2, 3))
    print(f'{tag:40s}', ['%.2e' % float(v) for v in e], flush=True)
run('all tensor-core (default)')
S.ops.mha = lambda *a, **k: orig_mha(*a, **{**k, 'exact': True})
run('attention exact, convs tensor-core')
S.ops.mha = orig_mha
S.ops.conv2d = lambda *a, **k: orig_conv(*a, **{**k, 'exact': True})
run('convs exact, attention tensor-core')
def conv_sel(pred):
    def f(x, cw, *a, **k):
        if pred(x, cw, k): k = {**k, 'exact': True}
        return orig_conv(x, cw, *a, **k)
    return f
S.ops.conv2d = conv_sel(lambda x, cw, k: k.get('pre') is not None)
run('convs with GN prologue exact')
S.ops.conv2d = conv_sel(lambda x, cw, k: k.get('pre') is None)
run('convs without GN prologue exact')
S.ops.conv2d = conv_sel(lambda x, cw, k: cw.kh == 1)
run('1x1 convs / linears exact')
S.ops.conv2d = conv_sel(lambda x, cw, k: x.shape[1] >= 128 and cw.kh == 3)
run('3x3 convs at >=128^2 exact')
S.ops.conv2d = conv_sel(lambda x, cw, k: x.shape[1] < 128 and cw.kh == 3)
run('3x3 convs below 128^2 exact')
S.ops.conv2d = conv_sel(lambda x, cw, k: cw.Cout > 128)
run('convs with Cout > 128 (unfused 3 MMAs) exact')
S.ops.conv2d = orig_conv
S.ops.TC_VARIANT = 256
run('no accumulation-bias correction')
S.ops.TC_VARIANT = 0
