import sys, time
sys.path.insert(0, '/root/repo')
import torch
import semantic_search_kd_b200 as pkg
dev = torch.device('cuda', 0)
n, d = 1_105_228, 384
g = torch.Generator(device=dev); g.manual_seed(5)
X = torch.randn((n, d), generator=g, device=dev); X = X / X.norm(dim=1, keepdim=True)
q = torch.randn((1, d), generator=g, device=dev); q = (q / q.norm(dim=1, keepdim=True)).contiguous()
sc = X.to(torch.bfloat16).float() @ q[0]
for name, Xc in (("random", X), ("ascending", X[torch.argsort(sc)].contiguous()), ("descending", X[torch.argsort(sc, descending=True)].contiguous()),
                 ("block-sorted (1024-row blocks ascending)", X[torch.argsort(sc).view(-1)[torch.randperm(n // 1024 * 1024, device=dev).view(-1, 1024).sort(dim=1).values.view(-1)]].contiguous())):
    for casc in (1, 0):
        idx = pkg.FlatIPIndex(d, metric='inner_product', device=0)
        idx.set_option("cascade", casc)
        idx.add(Xc)
        for k in (10,):
            idx.search_device(q, k); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): s, i = idx.search_device(q, k)
            e1.record(); torch.cuda.synchronize()
            print(f"{name:45s} cascade={casc} k={k}: {e0.elapsed_time(e1)/5*1e3:9.1f} us per search", flush=True)
        idx.close()
