import sys, time
sys.path.insert(0, '/root/repo')
import torch, numpy as np
import semantic_search_kd_b200 as pkg
dev = torch.device('cuda', 0)
idx = pkg.FlatIPIndex(384, metric='inner_product', device=0)
x = torch.randn(2000, 384, device=dev); x = x / x.norm(dim=1, keepdim=True)
idx.add(x)
Q = torch.randn(1024, 384, device=dev); Q = (Q / Q.norm(dim=1, keepdim=True)).contiguous()
out = (torch.empty((1, 10), dtype=torch.float32, device=dev), torch.empty((1, 10), dtype=torch.int64, device=dev))
for name, fn in (("fresh outputs", lambda i: idx.search_device(Q[i % 1024:i % 1024 + 1], 10)),
                 ("fresh outputs, stable", lambda i: idx.search_device(Q[i % 1024:i % 1024 + 1], 10, stable_queries=True)),
                 ("out= reused", lambda i: idx.search_device(Q[i % 1024:i % 1024 + 1], 10, out=out))):
    for i in range(200): fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(5000): fn(i)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name}: host enqueue {1e6 * (t1 - t0) / 5000:.1f} us/call, with drain {1e6 * (t2 - t0) / 5000:.1f} us/call")
