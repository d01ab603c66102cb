"""Round 2: what vertex ordering buys the CUDA-core SpMM.

Orderings of the same graph (rows AND columns relabelled, CSR re-canonicalised, capi.relabel):
  random   the generator's seeded random ids (the BASELINE shapes as benchmarked)
  degree   descending degree (cheap: one sort)
  loa      hcspmm_loa_reorder, bit-exact with the reference's LOI.cpp:660-805 (sequential greedy, one persistent CTA)
For each: SpMM ms (shipped selector, per-graph aux), the number of condensed 8-column blocks (sum of blockPartition:
fewer = more column sharing inside 16-row windows = what LOA optimises) and the LOA time.
--products-loa runs LOA on the products shape in a child process under a time limit."""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hc-spmm_b200")]
import torch
from hcspmm import capi, graphs

dev = torch.device("cuda", 0)


def t(fn, n=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def measure(rp, ci, dim):
    n = rp.numel() - 1
    bp, etc, etr, ht = capi.preprocess(ci, rp, "shipped")
    aux = capi.GraphAux(rp, ci, ht)
    x = torch.randn(n, dim, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    ms = t(lambda: capi.spmm_aux(x, rp, ci, bp, etc, etr, ht, aux))
    return {"spmm_ms": ms, "condensed_blocks": int(bp.sum()), "gflops": 2.0 * ci.numel() * dim / ms / 1e6}


def degree_perm(rp):
    deg = (rp[1:] - rp[:-1]).long()
    return torch.argsort(-deg, stable=True).to(torch.int32)      # perm[new] = old


def run(name, rp, ci, dim, do_loa):
    rec = {"graph": name, "n": rp.numel() - 1, "nnz": int(ci.numel()), "dim": dim, "orderings": {}}
    rec["orderings"]["random"] = measure(rp, ci, dim)
    rp2, ci2 = capi.relabel(rp, ci, degree_perm(rp))
    rec["orderings"]["degree"] = measure(rp2, ci2, dim)
    if do_loa:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        perm, sizes, nf = capi.loa_reorder(rp, ci)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        rp3, ci3 = capi.relabel(rp, ci, perm)
        r = measure(rp3, ci3, dim)
        r.update(loa_seconds=dt, blocks=int(sizes.numel()), full_blocks=int(nf))
        rec["orderings"]["loa"] = r
    print(json.dumps(rec), flush=True)
    return rec


if __name__ == "__main__":
    out = []
    if "--child-products-loa" in sys.argv:
        rp, ci, info = graphs.named("products", device=dev)
        out.append(run("products-shape", rp, ci, 128, True))
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_locality_products_loa.json"), "w"), indent=1)
        sys.exit(0)
    rp, ci = graphs.rmat(320000, 8000000, seed=5, device=dev)
    out.append(run("rmat 320k / 8M (random ids)", rp, ci, 128, True))
    rp, ci = graphs.sbm_dense_windows(131072, community=512, p_in=0.1, extra=8, seed=4, device=dev)
    perm = torch.randperm(131072, device=dev, generator=torch.Generator(device=dev).manual_seed(2)).to(torch.int32)
    rp, ci = capi.relabel(rp, ci, perm)                  # hide the communities: can LOA find them again?
    out.append(run("sbm 131k communities of 512 (ids shuffled)", rp, ci, 128, True))
    rp, ci, info = graphs.named("products", device=dev)
    out.append(run("products-shape", rp, ci, 128, False))
    del rp, ci
    torch.cuda.empty_cache()
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_locality.json"), "w"), indent=1)
    if "--products-loa" in sys.argv:
        limit = int(sys.argv[sys.argv.index("--products-loa") + 1])
        t0 = time.perf_counter()
        try:
            subprocess.run([sys.executable, __file__, "--child-products-loa"], timeout=limit, check=True)
        except subprocess.TimeoutExpired:
            print(json.dumps({"graph": "products-shape", "loa": f"not finished within {limit} s"}), flush=True)
        print(f"products LOA leg: {time.perf_counter() - t0:.1f} s")
