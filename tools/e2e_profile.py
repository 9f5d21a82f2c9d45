"""Where does the public-API step (HandNet.forward on host frames) spend its time beyond the kernels?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
import torch
import bench
from hn_b200 import runtime

dev = torch.device("cuda", 0)
net = bench.build_net(dev)
B = 8
rgb, depth = bench.synthetic_frames(1000, B)
rgb_pin, depth_pin = rgb.pin_memory(), depth.pin_memory()
def sync(): torch.cuda.synchronize()
with torch.inference_mode():
    imgs = rgb_pin.to(dev); dpt = depth_pin.to(dev)
    for _ in range(5): net(list(imgs.unbind(0)), depth_images=dpt)
    sync()
    def t(f, n=20):
        sync(); t0 = time.perf_counter()
        for _ in range(n): f()
        sync(); return (time.perf_counter() - t0) / n * 1e3
    print("H2D rgb+depth (pinned, 39 MB)      %.3f ms" % t(lambda: (rgb_pin.to(dev, non_blocking=True), depth_pin.to(dev, non_blocking=True))))
    step = net._steps[next(iter(net._steps))]
    print("graph replay only                  %.3f ms" % t(lambda: step.run()))
    print("weights_token                      %.3f ms" % t(lambda: runtime.weights_token(net)))
    lst = list(imgs.unbind(0))
    print("foreach_copy + depth copy          %.3f ms" % t(lambda: (torch._foreach_copy_(step.images, lst), step.depth.copy_(dpt))))
    print("full forward (device inputs)       %.3f ms" % t(lambda: net(lst, depth_images=dpt)))
    def full():
        a = rgb_pin.to(dev, non_blocking=True); b = depth_pin.to(dev, non_blocking=True)
        return net(list(a.unbind(0)), depth_images=b)
    print("H2D + forward, serial              %.3f ms" % t(full))
