"""Dev timing: single-call latency (host wall clock, synchronised) of generation from device-resident inputs, per batch size --
the serving / eval case (is_train=False).  Plain calls vs the CUDA-graph replay of pipeline.GraphedGenerator."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bevgen_b200 as bg

with_map = len(sys.argv) > 1 and sys.argv[1] == "map"
md = bg.synthetic.synthetic_map()
table = bg.MapTable([bg.maps.parse_map_dict(md)]) if with_map else None
for B in (1, 2, 4, 8):
    frames = [bg.synthetic.quantize_f16(bg.synthetic.frame_c3(i, map_data=md)) for i in range(B)]
    batch = bg.FrameBatch.pack(frames, 10, map_ids=[0] * B if with_map else None, points_dtype="av2")
    gen = bg.BevGenerator(table)
    dv = gen.upload(batch)
    ol = torch.empty((B, 290, 400, 720), device="cuda")
    om = torch.empty((B, 9, 400, 720), device="cuda") if with_map else None
    def plain():
        gen.generate(batch, dev=dv, with_map=with_map, out_lidar=ol, out_map=om)
    def plain_h2d():
        gen.generate(batch, with_map=with_map, out_lidar=ol, out_map=om)
    res = {}
    variants = {"plain (inputs resident)": plain, "plain (pinned host inputs)": plain_h2d}
    if hasattr(bg, "GraphedGenerator"):
        gg = bg.GraphedGenerator(table, num_frames=B, num_sweeps=10, max_points=batch.num_points, with_map=with_map)
        variants["graph replay (pinned host inputs)"] = lambda: gg.run(batch)
    for name, fn in variants.items():
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        n = 50
        t0 = time.perf_counter()
        for _ in range(n):
            fn(); torch.cuda.synchronize()
        res[name] = (time.perf_counter() - t0) / n * 1e6
    print(f"B {B}" + ("  +map" if with_map else "") + ":  " + "   ".join(f"{k} {v:7.1f} us/call ({v / B:6.1f} us/frame)" for k, v in res.items()), flush=True)
