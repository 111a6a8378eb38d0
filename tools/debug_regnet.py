import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adamvs_b200 import ops, synth
C, D, h, w, up = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
B = 1
i = {32: 0, 16: 1, 8: 2}[C]
sd = synth.fill_state_dict(synth.state_dict_shapes(8), 21)
p = f"DepthNet.{i}.reg_fuse"
names = {"conv1_w": ".conv1.conv.weight", "gates1_w": ".conv_gru1.conv_gates.0.weight",
         "gates1_b": ".conv_gru1.conv_gates.0.bias", "cand1_w": ".conv_gru1.convc.0.weight",
         "cand1_b": ".conv_gru1.convc.0.bias", "conv2_w": ".conv2.conv.weight",
         "gates2_w": ".conv_gru2.conv_gates.0.weight", "gates2_b": ".conv_gru2.conv_gates.0.bias",
         "cand2_w": ".conv_gru2.convc.0.weight", "cand2_b": ".conv_gru2.convc.0.bias",
         "up1_w": ".upconv1.weight", "up1_b": ".upconv1.bias", "out_w": ".upconv2d.weight", "out_b": ".upconv2d.bias"}
wd = {k: sd[p + v].cuda() for k, v in names.items()}
vol = torch.randn(B, C, D, h, w).cuda()
cur = (600 + 10 * torch.randn(B, h, w)).cuda()
hyp = ops.Hyp(ops.HYP_PER_PIXEL, cur, torch.tensor([3.3]).cuda())
d, c = ops.regnet_red(vol, wd, hyp, bool(up), ops.PROB_SOFTMAX)
torch.cuda.synchronize()
print("ok", float(d.mean()), float(c.mean()))
