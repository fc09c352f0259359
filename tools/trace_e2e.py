import sys, os, torch
sys.path.insert(0, "/root/repo")
import combo_avs_b200 as pkg
from combo_avs_b200.host_api import HostPipeline
from bench import S4_SHAPES, synth_inputs, LQ, M, D
dev = torch.device("cuda:0")
shapes = torch.tensor(S4_SHAPES, dtype=torch.long, device=dev)
start = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
hin = synth_inputs(torch, 40, 3, torch.float32, torch.float32, None, pin=True)
hout = [torch.empty((40, LQ, M * D)).pin_memory(), torch.empty_like(hin[0]).pin_memory(), torch.empty_like(hin[1]).pin_memory(), torch.empty_like(hin[2]).pin_memory()]
pipe = HostPipeline(dev, chunk_frames=int(sys.argv[1]), slots=3)
for i in range(3):
    print("call", i, file=sys.stderr)
    pipe.forward_backward(hin[0], shapes, start, hin[1], hin[2], hin[3], *hout, 128)
