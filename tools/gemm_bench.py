"""GEMM micro-benchmark on the BERT shapes: rvn tcgen05 kernel vs cuBLAS (torch.matmul), CUDA-event timed,
L2 flushed between iterations.  Writes gpurun_out/gemm_bench.json; `--once` runs each kernel a few times (for ncu)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ravnest_b200.ops import _ext  # noqa: E402

C = _ext.load()
dev = 'cuda'
once = '--once' in sys.argv
SHAPES = [  # (name, M, N, K, a_mn, b_mn)
    ('qkv_fwd', 4096, 2304, 768, False, False), ('proj_fwd', 4096, 768, 768, False, False), ('ffn1_fwd', 4096, 3072, 768, False, False),
    ('ffn2_fwd', 4096, 768, 3072, False, False), ('vocab_fwd', 4096, 30528, 768, False, False),
    ('ffn1_dgrad', 4096, 768, 3072, False, True), ('ffn1_wgrad', 3072, 768, 4096, True, True),
    ('vocab_wgrad', 30528, 768, 4096, True, True), ('square_8k', 8192, 8192, 8192, False, False),
]
SHAPES += [('proj_dgrad', 4096, 768, 768, False, True), ('qkv_dgrad', 4096, 768, 2304, False, True), ('ffn2_dgrad', 4096, 3072, 768, False, True),
           ('proj_wgrad', 768, 768, 4096, True, True), ('qkv_wgrad', 2304, 768, 4096, True, True), ('ffn2_wgrad', 768, 3072, 4096, True, True)]
sweep = '--sweep' in sys.argv
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
results = []


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    times = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e))
    return sorted(times)[len(times) // 2]


if sweep:       # every (block_n, split_k) the kernel offers, per shape; wgrad shapes in accumulate (TMA reduce-add) mode
    for name, M, N, K, amn, bmn in SHAPES:
        A = (torch.randn((K, M) if amn else (M, K), device=dev) * 0.1).to(torch.bfloat16)
        Bm = (torch.randn((K, N) if bmn else (N, K), device=dev) * 0.1).to(torch.bfloat16)
        wgrad = amn and bmn
        G = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
        row = {'name': name, 'M': M, 'N': N, 'K': K, 'cublas_ms': timed(lambda: (A.t() if amn else A) @ (Bm if bmn else Bm.t()))}
        for bn in (128, 192, 256):
            for sk in ((1, 2, 3, 4, 6, 8) if wgrad else (1,)):
                if wgrad:
                    fn = lambda: C.gemm(A, amn, Bm, bmn, None, 0, G, False, bn, 2, G, None, None, 0, sk)
                else:
                    fn = lambda: C.gemm(A, amn, Bm, bmn, None, 0, None, False, bn)
                row['bn{}_sk{}'.format(bn, sk)] = round(timed(fn, 7) * 1e3, 2)
        for bn in (128, 256):                   # CTA-pair kernel (cta_group::2), plain epilogue
            row['2cta_bn{}'.format(bn)] = round(timed(lambda: C.gemm_2cta(A, amn, Bm, bmn, None, 0, bn), 7) * 1e3, 2)
        if wgrad:
            row['auto_us'] = round(timed(lambda: C.gemm(A, amn, Bm, bmn, None, 0, G, False, 0, 2, G)) * 1e3, 2)
        else:
            row['auto_us'] = round(timed(lambda: C.gemm(A, amn, Bm, bmn, None, 0, None, False, 0)) * 1e3, 2)
        row['cublas_us'] = round(row.pop('cublas_ms') * 1e3, 2)
        best = min((v, k) for k, v in row.items() if k.startswith('bn'))
        row['best'] = best[1]
        print(json.dumps(row))
        results.append(row)
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(results, open('gpurun_out/gemm_sweep.json', 'w'), indent=1)
    sys.exit(0)

for name, M, N, K, amn, bmn in SHAPES:
    A = (torch.randn((K, M) if amn else (M, K), device=dev) * 0.1).to(torch.bfloat16)
    Bm = (torch.randn((K, N) if bmn else (N, K), device=dev) * 0.1).to(torch.bfloat16)

    def ours():
        return C.gemm(A, amn, Bm, bmn, None, 0, None, False, 0)[0]

    def cublas():
        a = A.t() if amn else A
        b = Bm if bmn else Bm.t()
        return a @ b

    row = {'name': name, 'M': M, 'N': N, 'K': K}
    for label, fn in (('rvn', ours), ('cublas', cublas)):
        for _ in range(1 if once else 3):
            fn()
        if once:
            continue
        times = []
        for _ in range(10):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            times.append(s.elapsed_time(e))
        t = sorted(times)[len(times) // 2]
        row[label + '_ms'] = t
        row[label + '_tflops'] = 2.0 * M * N * K / t / 1e9
    results.append(row)
    print(row)
if not once:
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(results, open('gpurun_out/gemm_bench.json', 'w'), indent=1)
