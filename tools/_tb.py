import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from _cases import biased, english
huf = importlib.import_module("huffman-avx512_b200")
codec = huf.BlockCodec(32, 131072)
for name, data in [("biased128k", biased(131072, seed=1)), ("biased16k", biased(16384, seed=1)), ("english128k", english(131072, seed=1)), ("uniform128k", np.random.default_rng(0).integers(0, 256, 131072, dtype=np.uint8).tobytes())]:
    raw = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    hist = codec.histogram(raw)
    tab = codec.build_table(hist)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): codec.build_table(hist, out=tab)
    e1.record(); torch.cuda.synchronize()
    print(name, "distinct", int((hist > 0).sum()), f"{e0.elapsed_time(e1) / 200 * 1e3:.1f} us per k_build_table launch")
