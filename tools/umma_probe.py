"""GPU probe of tqb_apply_blocks (tcgen05 dense 6-qubit block): numerics vs a float64 torch contraction for
several block placements (normal and transposed write-back) and timing at n = 28.
Run on the GPU box: python tools/umma_probe.py [--n 20] [--time-n 28]"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def tf32_rna(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def a_planes(U):
    """U: complex128 [64,64] -> float32 [2,128,64] (TF32 hi plane, TF32 lo plane), row 2i+w = (Re, Im)[w] U[i]"""
    A = torch.empty(128, 64, dtype=torch.float64)
    A[0::2] = U.real
    A[1::2] = U.imag
    A = A.to(torch.float32)
    hi = tf32_rna(A)
    lo = tf32_rna(A - hi)
    return torch.stack([hi, lo]).contiguous()


def reference(state, U, kpos, n, transpose):
    """state complex128 [2^n]; U[i][j] with pattern bit b <-> index bit kpos[b]"""
    t = state.reshape([2] * n)
    axes = [n - 1 - kpos[b] for b in reversed(range(6))]  # most significant pattern bit first
    rest = [a for a in range(n) if a not in axes]
    perm = axes + rest
    v = t.permute(perm).reshape(64, -1)
    o = (U @ v).reshape([2] * n)
    inv = [0] * n
    for i, a in enumerate(perm):
        inv[a] = i
    o = o.permute(inv)
    if transpose:
        for b in range(6):
            o = o.transpose(n - 1 - b, n - 1 - kpos[b])
    return o.reshape(-1)


def rotated(vec, kpos, n, transpose):
    """the layout a rotating block leaves behind: block bits -> index bits 6-11 (pattern order; 0-5 when transposing, the
    old bits 0-5 then at 6-11), every other bit >= 6 compacted above them in ascending order.  vec: reference(...,
    transpose=False) output, i.e. the result in the SOURCE layout."""
    src_of = {}           # destination bit -> source bit
    for b in range(6):
        src_of[b if transpose else 6 + b] = kpos[b]
        src_of[6 + b if transpose else b] = b
    r = 12
    for bit in range(6, n):
        if bit not in kpos:
            src_of[r] = bit
            r += 1
    t = vec.reshape([2] * n)
    # destination axis a (bit n-1-a) takes source axis n-1-src_of[n-1-a]
    return t.permute([n - 1 - src_of[n - 1 - a] for a in range(n)]).reshape(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20)
    ap.add_argument("--time-n", type=int, default=28)
    ap.add_argument("--out", default="gpurun_out/umma_probe.json")
    args = ap.parse_args()
    lib = ctypes.CDLL(os.path.join(ROOT, "torchquantum_b200", "lib", "libtqb200.so"))
    lib.tqb_last_error.restype = ctypes.c_char_p
    fn = lib.tqb_apply_blocks
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p,
                   ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    fn.restype = ctypes.c_int
    renorm = [1]
    dev = "cuda:0"
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(1)
    M = torch.randn(64, 64, dtype=torch.float64, generator=g) + 1j * torch.randn(64, 64, dtype=torch.float64, generator=g)
    U, _ = torch.linalg.qr(M)
    planes = a_planes(U).to(dev)
    Ud = U.to(dev)
    n = args.n
    results = {"cases": []}
    placements = [list(range(6, 12)), list(range(10, 16)), [6, 7, 8, 12, 13, 14], [6, 7, 9, 10, n - 2, n - 1],
                  list(range(n - 6, n))]

    def run(src, dst, nq, bsz, kpos, tr, rot=0):
        hdr = [0, 0, tr | (2 * rot), 0] + list(kpos) + [0, 0]   # csrc/tqb_program.h: TQB_KB_*
        arr = (ctypes.c_int32 * 12)(*hdr)
        rc = fn(src.data_ptr(), dst.data_ptr(), None, nq, bsz, planes.data_ptr(), ctypes.addressof(arr), 1, renorm[0],
                torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            raise RuntimeError(lib.tqb_last_error().decode())

    for bsz in ((1, 2) if n >= 12 else ()):
        for kpos in placements:
            for tr in (0, 1):
                st = torch.randn(bsz, 2 ** n, 2, device=dev, dtype=torch.float32)
                st = st / st.pow(2).sum((1, 2), keepdim=True).sqrt()
                out = torch.empty_like(st)
                try:
                    run(st, out, n, bsz, kpos, tr)
                    torch.cuda.synchronize()
                except Exception as e:  # noqa: BLE001
                    results["cases"].append({"kpos": kpos, "tr": tr, "bsz": bsz, "error": str(e)})
                    print("ERROR", kpos, tr, e, flush=True)
                    continue
                worst = 0.0
                rel = 0.0
                for b in range(bsz):
                    ref = reference(torch.view_as_complex(st[b].double()), Ud, kpos, n, tr)
                    got = torch.view_as_complex(out[b].double())
                    worst = max(worst, (got - ref).abs().max().item() / ref.abs().max().item())
                    rel = max(rel, ((got - ref).norm() / ref.norm()).item())
                results["cases"].append({"kpos": kpos, "tr": tr, "bsz": bsz, "max_rel": worst, "rel_l2": rel})
                print(kpos, "tr", tr, "bsz", bsz, "max_rel %.3e rel_l2 %.3e" % (worst, rel), flush=True)
                # in place
                st2 = st.clone()
                run(st2, st2, n, bsz, kpos, tr)
                torch.cuda.synchronize()
                same = torch.equal(st2, out)
                if not same:
                    print("  in-place differs from out-of-place!", (st2 - out).abs().max().item(), flush=True)
                results["cases"][-1]["inplace_equal"] = bool(same)

    # rotating blocks (gathering reads, contiguous writes) incl. unordered block bits and enumerated pieces
    rot_cases = [list(range(6, 12)), list(range(n - 6, n)), [n - 1, n - 2, n - 3, 8, 7, 6], [6, 8, 10, 12, n - 3, n - 1],
                 [9, 10, 13, 14, 7, n - 1], [n - 2, 6, n - 4, 8, 11, 10], [7, 8, 9, n - 3, n - 2, n - 1]]
    for bsz in ((1, 2) if n >= 12 else ()):
        for kpos in rot_cases:
            for tr in (0, 1):
                st = torch.randn(bsz, 2 ** n, 2, device=dev, dtype=torch.float32)
                st = st / st.pow(2).sum((1, 2), keepdim=True).sqrt()
                out = torch.zeros_like(st)
                try:
                    run(st, out, n, bsz, kpos, tr, rot=1)
                    torch.cuda.synchronize()
                except Exception as e:  # noqa: BLE001
                    results["cases"].append({"kpos": kpos, "tr": tr, "bsz": bsz, "rotate": 1, "error": str(e)})
                    print("ERROR rotate", kpos, tr, e, flush=True)
                    continue
                rel = 0.0
                for b in range(bsz):
                    ref = rotated(reference(torch.view_as_complex(st[b].double()), Ud, kpos, n, 0), kpos, n, tr)
                    got = torch.view_as_complex(out[b].double())
                    rel = max(rel, ((got - ref).norm() / ref.norm()).item())
                results["cases"].append({"kpos": kpos, "tr": tr, "bsz": bsz, "rotate": 1, "rel_l2": rel})
                print("rotate", kpos, "tr", tr, "bsz", bsz, "rel_l2 %.3e" % rel, flush=True)
        # unordered / enumerated block bits, in place
        for kpos in rot_cases[2:]:
            st = torch.randn(bsz, 2 ** n, 2, device=dev, dtype=torch.float32)
            st = st / st.pow(2).sum((1, 2), keepdim=True).sqrt()
            ref = torch.stack([reference(torch.view_as_complex(st[b].double()), Ud, kpos, n, 0) for b in range(bsz)])
            run(st, st, n, bsz, kpos, 0)
            torch.cuda.synchronize()
            rel = ((torch.view_as_complex(st.double()) - ref).norm() / ref.norm()).item()
            results["cases"].append({"kpos": kpos, "tr": 0, "bsz": bsz, "inplace_unordered": 1, "rel_l2": rel})
            print("in place, unordered", kpos, "bsz", bsz, "rel_l2 %.3e" % rel, flush=True)

    # norm drift over 64 applications of the same block, with and without the per-tile renormalisation
    for rn in ((0, 1) if n >= 12 else ()):
        renorm[0] = rn
        st = torch.randn(1, 2 ** n, 2, device=dev, dtype=torch.float32)
        st /= st.double().norm().float()
        n0 = st.double().norm().item()
        for _ in range(64):
            run(st, st, n, 1, list(range(8, 14)), 0)
        torch.cuda.synchronize()
        results["norm_drift_64_renorm%d" % rn] = st.double().norm().item() / n0 - 1.0
        print("norm drift after 64 blocks, renorm", rn, results["norm_drift_64_renorm%d" % rn], flush=True)
    renorm[0] = 1
    if args.time_n < 28:     # (the timing placements below are written for n >= 28; --time-n 0: numerics only)
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)
        print("numerics only:", len(results["cases"]), "cases, worst rel_l2",
              max(c.get("rel_l2", 0.0) for c in results["cases"]), "errors", sum(1 for c in results["cases"] if "error" in c))
        return
    # timing
    nt = args.time_n
    st = torch.randn(1, 2 ** nt, 2, device=dev, dtype=torch.float32)
    st /= st.norm()
    timing = {}
    for name, kpos, tr in (("low_block", list(range(6, 12)), 0), ("mid_block", list(range(14, 20)), 0),
                           ("top_block", list(range(nt - 6, nt)), 0), ("mid_block_T", list(range(14, 20)), 1),
                           ("three_runs", [6, 7, 12, 13, 20, 21], 0)):
        for _ in range(3):
            run(st, st, nt, 1, kpos, tr)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
        ev[0].record()
        for k in range(10):
            run(st, st, nt, 1, kpos, tr)
            ev[k + 1].record()
        torch.cuda.synchronize()
        ms = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(10))
        gbs = 16 * 2 ** nt / (ms[len(ms) // 2] * 1e-3) / 1e9
        timing[name] = {"ms_median": ms[len(ms) // 2], "ms_min": ms[0], "GBps": gbs}
        print(name, timing[name], flush=True)
    # rotating passes: gathering reads from any placement, contiguous 32 KiB writes into the second buffer
    bufs = (st, torch.empty_like(st))
    for name, kpos, tr in (("rot_low", list(range(6, 12)), 0), ("rot_mid", list(range(14, 20)), 0),
                           ("rot_19_24", list(range(19, 25)), 0), ("rot_top", list(range(nt - 6, nt)), 0),
                           ("rot_top_T", list(range(nt - 6, nt)), 1), ("rot_12_13_top4", [12, 13] + list(range(nt - 4, nt)), 0),
                           ("rot_3runs", [6, 7, 14, 15, 22, 23], 0), ("rot_4runs_ext1", [8, 9, 13, 14, 20, nt - 1], 0),
                           ("rot_6runs_ext3", [7, 9, 12, 15, 20, 25], 0), ("rot_unordered", [25, 24, 9, 8, 17, 16], 1)):
        a, b = bufs
        for _ in range(3):
            run(a, b, nt, 1, kpos, tr, rot=1)
            a, b = b, a
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
        ev[0].record()
        for k in range(10):
            run(a, b, nt, 1, kpos, tr, rot=1)
            a, b = b, a
            ev[k + 1].record()
        torch.cuda.synchronize()
        ms = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(10))
        gbs = 16 * 2 ** nt / (ms[len(ms) // 2] * 1e-3) / 1e9
        timing[name] = {"ms_median": ms[len(ms) // 2], "ms_min": ms[0], "GBps": gbs}
        print(name, timing[name], flush=True)
        bufs = (a, b)
    st = bufs[0]
    results["timing_n%d" % nt] = timing
    results["norm_after"] = st.double().norm().item()
    print("norm after", results["norm_after"])
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
