import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import train as OT
from uni3detr_b200 import synth
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_train import _gts, rel
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
model, cfg = synth.build_model("sunrgbd", seed=0)
for m in model.modules():
    if isinstance(m, torch.nn.Dropout): m.p = 0.0
    if isinstance(m, torch.nn.MultiheadAttention): m.dropout = 0.0
scenes = [synth.make_scene("sunrgbd", 0, n_points=2500), synth.make_scene("sunrgbd", 1, n_points=1800)]
rng = np.random.default_rng(0)
pcr = cfg["pts_voxel_layer"]["point_cloud_range"]
gts, gls = zip(*[_gts(rng, n, pcr, 10) for n in (4, 6)])
sd = {k: v.detach().clone().float() for k, v in model.state_dict().items()}
names = [k for k, p in model.named_parameters() if p.requires_grad]
for k in names: sd[k].requires_grad_()
want = OT.forward_train(sd, cfg, scenes, list(gts), list(gls))
sum(want.values()).backward()
model = model.to("cuda").train()
got = model(return_loss=True, points=[torch.from_numpy(s).cuda() for s in scenes], img_metas=[{}, {}],
            gt_bboxes_3d=[t.cuda() for t in gts], gt_labels_3d=[t.cuda() for t in gls])
for k in want: print("loss %-20s %.6f %.6f" % (k, float(got[k]), float(want[k])))
sum(got.values()).backward()
params = dict(model.named_parameters())
for k in names:
    gw, gp = sd[k].grad, params[k].grad
    print("%-80s rel %.2e  |oracle|max %.3e  cos %.6f" % (k, rel(gp, gw), float(gw.abs().max()),
          float(torch.nn.functional.cosine_similarity(gp.detach().cpu().reshape(1, -1), gw.reshape(1, -1)))))
