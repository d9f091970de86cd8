"""Config 5 of BASELINE.json end to end: "BEV generation + flip/rotate/scale/dropout augmentation feeding IntentNetViT forward
at batch 16 per GPU".

The generator side is the product (pinned host frames -> H2D -> fused generation + augmentation -> device-resident
[B,290,400,720] / [B,9,400,720]).  The consumer is a SINK, not a parity target: the reference's IntentNetViT needs `timm`
(model_vit.py:4,64,71), which is not installed here, so a torch-only stand-in with the same tensor interface is used --
two ViT-S/8 streams (patch 8, dim 384, 12 blocks, 6 heads, in_chans 290 / 9, 50 x 90 = 4 500 tokens each; model_vit.py:38-83),
LayerNorm + Linear + GELU adapters to 192 channels (model_vit.py:82-83), concat to [B,384,50,90], a small conv fusion block
and 3x3 conv heads (heads.py:6-42) -- run under bf16 autocast.  What is measured:

    consumer alone          forward on a fixed device-resident batch
    generator alone         the bench's pipelined e2e loop for config 5
    fed                     generation of batch i+1 (generator streams) overlapped with the forward of batch i

usage: python tools/feed_vit_consumer.py [batch=16] [steps=6] [points_format=av2]
"""
import random
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

import bevgen_b200 as bg


class Block(nn.Module):
    def __init__(self, dim=384, heads=6):
        super().__init__()
        self.n1, self.n2 = nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.qkv, self.proj = nn.Linear(dim, 3 * dim), nn.Linear(dim, dim)
        self.fc1, self.fc2 = nn.Linear(dim, 4 * dim), nn.Linear(4 * dim, dim)
        self.heads = heads

    def forward(self, x):
        B, N, C = x.shape
        q, k, v = self.qkv(self.n1(x)).view(B, N, 3, self.heads, C // self.heads).permute(2, 0, 3, 1, 4)
        x = x + self.proj(F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, C))
        return x + self.fc2(F.gelu(self.fc1(self.n2(x))))


class ViTStream(nn.Module):
    def __init__(self, in_chans, dim=384, depth=12):
        super().__init__()
        self.patch = nn.Conv2d(in_chans, dim, 8, 8)
        self.cls = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos = nn.Parameter(torch.zeros(1, 4501, dim))
        self.blocks = nn.Sequential(*[Block(dim) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim)
        self.adapter = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, 192), nn.GELU())

    def forward(self, x):
        t = self.patch(x).flatten(2).transpose(1, 2)
        t = torch.cat([self.cls.expand(t.shape[0], -1, -1), t], 1) + self.pos
        t = self.adapter(self.norm(self.blocks(t))[:, 1:])
        return t.transpose(1, 2).reshape(x.shape[0], 192, 50, 90)


class Sink(nn.Module):
    def __init__(self):
        super().__init__()
        self.lidar, self.map = ViTStream(290), ViTStream(9)
        self.fuse = nn.Sequential(nn.Conv2d(384, 512, 3, padding=1), nn.BatchNorm2d(512), nn.ReLU(),
                                  nn.Conv2d(512, 512, 3, padding=1), nn.BatchNorm2d(512), nn.ReLU())
        self.det, self.intent = nn.Conv2d(512, 5 * 7, 3, padding=1), nn.Conv2d(512, 5 * 8, 3, padding=1)

    def forward(self, lidar_bev, map_bev):
        f = self.fuse(torch.cat([self.lidar(lidar_bev), self.map(map_bev)], 1))
        return self.det(f), self.intent(f)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    fmt = sys.argv[3] if len(sys.argv) > 3 else "av2"
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = Sink().to(dev).eval()
    md = bg.synthetic.synthetic_map()
    base = [bg.synthetic.frame_c3(i, map_data=md) for i in range(2)]
    if fmt != "f32":
        base = [bg.synthetic.quantize_f16(f) for f in base]
    frames = [base[i % 2] for i in range(B)]
    aug = []
    for i in range(B):
        random.seed(5000 + i)
        aug.append(bg.draw_aug_params(random))
    table = bg.MapTable([bg.maps.parse_map_dict(md)])
    batch = bg.FrameBatch.pack(frames, 10, map_ids=[0] * B, points_dtype={"av2": "av2", "f16": np.float16, "f32": np.float32}[fmt])
    gen = bg.BevGenerator(table)
    pf = bg.PrefetchingGenerator(gen)
    bufs = [(torch.empty((B, 290, 400, 720), device=dev), torch.empty((B, 9, 400, 720), device=dev)) for _ in range(2)]
    gen_stream, fwd_stream = torch.cuda.Stream(), torch.cuda.Stream()

    def forward(l, m):
        with torch.inference_mode(), torch.autocast("cuda", dtype=torch.bfloat16):
            return model(l, m)

    def timed(fn, n):
        fn(2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn(n)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n

    r0 = gen.generate(batch, aug, out_lidar=bufs[0][0], out_map=bufs[0][1])
    torch.cuda.synchronize()

    def consumer_only(n):
        for _ in range(n):
            forward(r0["lidar_bev"], r0["map_bev"])

    def generator_only(n):
        # the host stays at most two steps ahead of the GPU (as bench.py's e2e loop does): the staging tensors of a finished
        # step go back to the allocator instead of forcing a fresh cudaMalloc per step
        h = pf.submit(batch)
        done = []
        for i in range(n):
            if len(done) >= 2:
                done.pop(0).synchronize()
            cur, h = h, (pf.submit(batch) if i + 1 < n else None)
            pf.run(cur, aug, out_lidar=bufs[i % 2][0], out_map=bufs[i % 2][1])
            ev = torch.cuda.Event()
            ev.record()
            done.append(ev)

    def fed(n):
        # generation of batch i+1 on gen_stream while the forward of batch i runs on fwd_stream; two output buffers
        done_fwd = [None, None]
        h = pf.submit(batch)
        for i in range(n):
            if done_fwd[i % 2] is not None:
                done_fwd[i % 2].synchronize()                     # host at most two steps ahead (bounded staging memory)
            with torch.cuda.stream(gen_stream):
                if done_fwd[i % 2] is not None:
                    gen_stream.wait_event(done_fwd[i % 2])            # the buffer's previous consumer has finished
                cur, h = h, (pf.submit(batch) if i + 1 < n else None)
                r = pf.run(cur, aug, out_lidar=bufs[i % 2][0], out_map=bufs[i % 2][1])
                ready = torch.cuda.Event()
                ready.record(gen_stream)
            with torch.cuda.stream(fwd_stream):
                fwd_stream.wait_event(ready)
                forward(r["lidar_bev"], r["map_bev"])
                ev = torch.cuda.Event()
                ev.record(fwd_stream)
                done_fwd[i % 2] = ev

    tg = timed(generator_only, steps)
    tc = timed(consumer_only, steps)
    tf = timed(fed, steps)
    tg2 = timed(generator_only, steps)        # again after the forward passes (allocator / clock state differs)
    n_res = sum(1 for p in aug if p.needs_gather)
    print(f"config 5, batch {B} ({n_res} frames rotate/scale), points {fmt}, H2D {batch.h2d_bytes / 1e6:.0f} MB/step")
    print(f"  consumer alone (ViT-S/8 x 2 stand-in, bf16 autocast): {tc * 1e3:8.2f} ms/step  {B / tc:8.0f} frames/s")
    print(f"  generator alone (H2D + generation + augmentation):    {tg * 1e3:8.2f} ms/step  {B / tg:8.0f} frames/s")
    print(f"  generator alone, measured again after the forwards:   {tg2 * 1e3:8.2f} ms/step  {B / tg2:8.0f} frames/s")
    print(f"  fed (generation of batch i+1 under forward of i):     {tf * 1e3:8.2f} ms/step  {B / tf:8.0f} frames/s")
    print(f"  generation + augmentation add {1e3 * (tf - tc):.2f} ms ({100 * (tf - tc) / tc:.1f} %) to the consumer's step: the input pipeline is "
          f"{tc / tg:.0f} x faster than the forward it feeds (the reference's CPU loader: ~1 frame/s per process, i.e. ~{B:d} s per step)")


if __name__ == "__main__":
    main()
