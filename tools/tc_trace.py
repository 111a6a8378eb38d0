"""Role timeline of the tcgen05 conv kernel (GRU-1 gate convolution, block 0): needs a debug build,
    ADAMVS_TC_TRACE=1 python adamvs_b200/build.py --force
usage: python tools/tc_trace.py [--batch 8]"""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from adamvs_b200 import ops, synth
from tools.tc_regnet_check import NAMES


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--tf32", action="store_true")
    a = ap.parse_args()
    sd = synth.fill_state_dict(synth.state_dict_shapes(8), 21)
    C, h, w, up = 8, 384, 768, False
    B, D = a.batch, 2
    p = "DepthNet.2.reg_fuse"
    wd = {k: sd[p + v].cuda() for k, v in NAMES.items()}
    vol = torch.randn(B, C, D, h, w).cuda()
    cur = (600 + 10 * torch.randn(B, h, w)).cuda()
    hyp = ops.Hyp(ops.HYP_PER_PIXEL, cur, torch.tensor([3.3]).cuda())
    for _ in range(2):
        ops.regnet_red(vol, wd, hyp, up, ops.PROB_SOFTMAX, math=ops.MATH_TC_FP32)
    torch.cuda.synchronize()
    L = ops.lib()
    buf = np.zeros((4, 512, 4), dtype=np.int64)
    rc = L.adamvs_tc_trace_read(buf.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, rc
    conv, mma, epi, ex = buf
    t0 = min(int(conv[0, 0]), int(mma[0, 0]))
    n = int((mma[:, 2] > 0).sum())
    print(f"chunks traced: {n}")
    print("chunk | conv: top  got_stage  stored  fetched | mma: top  got_full  issued | (clk since start)")
    for g in range(min(n, 40)):
        c = [int(x) - t0 for x in conv[g]]
        m = [int(x) - t0 for x in mma[g][:3]]
        print(f"{g:4d} | {c[0]:8d} {c[1]:8d} {c[2]:8d} {c[3]:8d} | {m[0]:8d} {m[1]:8d} {m[2]:8d}")
    nt = int((epi[:, 2] > 0).sum())
    print("tile | epi: top  got_acc  done")
    for t in range(min(nt, 20)):
        e = [int(x) - t0 for x in epi[t][:3]]
        x = [int(v) - t0 for v in ex[t][:3]]
        print(f"{t:4d} | {e[0]:8d} {e[1]:8d} {e[2]:8d} | step0: tmem {x[0]:8d} math {x[1]:8d} stored {x[2]:8d}")
    if n > 8:
        print("steady state clk/chunk (mma issued):", (int(mma[n - 1, 2]) - int(mma[4, 2])) / (n - 5))
        d_slot = (conv[5:n, 3] - conv[5:n, 0]).mean(); d_stage = (conv[5:n, 1] - conv[5:n, 3]).mean(); d_store = (conv[5:n, 2] - conv[5:n, 1]).mean()
        gap = (conv[6:n, 0] - conv[5:n - 1, 2]).mean()
        print(f"converter: wait TMA slot {d_slot:.0f}, wait operand stage free {d_stage:.0f}, split+store {d_store:.0f}, loop gap {gap:.0f}")
        first = mma[0:n:2]                       # first chunk of a tile carries the accumulator wait stamp (gates1: 2 chunks per tile)
        d_acc = (first[3:, 0] - first[3:, 3]).mean()
        print(f"mma thread: wait accumulator free (per tile) {d_acc:.0f}")
        m_wait = (mma[5:n, 1] - mma[5:n, 0]).mean(); m_issue = (mma[5:n, 2] - mma[5:n, 1]).mean()
        print(f"mma thread: wait full {m_wait:.0f}, issue {m_issue:.0f}")
        e_wait = (epi[2:nt, 1] - epi[2:nt, 0]).mean(); e_run = (epi[2:nt, 2] - epi[2:nt, 1]).mean()
        print(f"epilogue: wait acc {e_wait:.0f}, run {e_run:.0f}")


if __name__ == "__main__":
    main()
