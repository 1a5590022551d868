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
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
results = []
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
