"""Tiny launcher for `ncu` captures: runs ONE named case a few times so `ncu -k regex:<kernel> -c N` sees the shipped kernel
on a BASELINE shape.   python tools/ncu_targets.py gemm_f16|gemm_bf16|gemm_t|lut_2x8|lut_8x8|gemv_1x16|layer_1x16"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aqlm_b200.inference_kernels import cuda_kernel  # noqa: E402

DEV = "cuda:0"


def weights(fin, fout, K, nbits, dt):
    lo, hi = (-128, 128) if nbits <= 8 else (-32768, 32768)
    return (torch.randint(lo, hi, (fout, fin // 8, K), dtype=torch.int8 if nbits <= 8 else torch.int16, device=DEV),
            torch.randn((K, 2**nbits, 1, 8), dtype=dt, device=DEV),
            (0.75 + 0.5 * torch.rand((fout, 1, 1, 1), device=DEV)).to(dt))


def main():
    case = sys.argv[1]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    if case in ("gemm_f16", "gemm_bf16"):
        dt = torch.float16 if case == "gemm_f16" else torch.bfloat16
        w = weights(4096, 14336, 1, 16, dt)
        x = torch.randn((256, 4096), dtype=dt, device=DEV)
        for _ in range(reps):
            cuda_kernel.matmat_dequant(x, *w, None)
    elif case == "gemm_t":
        w = weights(4096, 14336, 1, 16, torch.float16)
        go = torch.randn((256, 14336), dtype=torch.float16, device=DEV)
        for _ in range(reps):
            cuda_kernel.matmat_dequant_transposed(go, *w, None)
    elif case in ("lut_2x8", "lut_8x8"):
        K = 2 if case == "lut_2x8" else 8
        w = weights(4096, 11008, K, 8, torch.float16)
        x = torch.randn((1, 4096), dtype=torch.float16, device=DEV)
        for _ in range(reps):
            cuda_kernel.matmat(x, *w, None)
    elif case == "gemv_1x16":
        w = weights(4096, 14336, 1, 16, torch.float16)
        x = torch.randn((1, 4096), dtype=torch.float16, device=DEV)
        for _ in range(reps):
            cuda_kernel.matmat(x, *w, None)
    elif case == "layer_1x16":  # the 4 grouped launches of one Llama-3-8B decoder layer, as bench.py runs them
        import aqlm_b200

        def lin(fin, fout):
            m = aqlm_b200.QuantizedLinear(fin, fout, 8, 1, 1, 16, bias=False, device=DEV, dtype=torch.float16)
            m.codes.data, m.codebooks.data, m.scales.data = weights(fin, fout, 1, 16, torch.float16)
            return m
        qkv = aqlm_b200.QuantizedLinearGroup([lin(4096, 4096), lin(4096, 1024), lin(4096, 1024)])
        o = lin(4096, 4096)
        gu = aqlm_b200.QuantizedLinearGroup([lin(4096, 14336), lin(4096, 14336)])
        down = lin(14336, 4096)
        x = torch.randn((1, 4096), dtype=torch.float16, device=DEV)
        xi = torch.randn((1, 14336), dtype=torch.float16, device=DEV)
        for _ in range(reps):
            qkv(x); o(x); gu(x); down(xi)
    elif case.startswith("shards_70b_n"):  # the 4 fused GEMV+exchange launches of one Llama-3-70B layer on 1/N-size shards
        import torch.distributed as dist

        import aqlm_b200
        from aqlm_b200.grouped import ShardedQuantizedLinearGroup
        from aqlm_b200.peer import PeerComm
        from aqlm_b200.sharded import ShardedQuantizedLinear

        n = int(case[len("shards_70b_n"):])
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
        dist.init_process_group("gloo", rank=0, world_size=1)
        comm = PeerComm(max_elems=4 * 28672)  # a one-rank communicator: the kernel pushes to itself, same DRAM traffic

        def lin(fin, fout):
            m = ShardedQuantizedLinear(fin // n, fout, 8, 1, 1, 16, bias=False, rank=0, world_size=1, device=DEV,
                                       dtype=torch.float16, peer_comm=comm)
            m.world_size = 2  # take the exchange path (the communicator itself has one rank)
            m.in_begin, m.in_end = 0, fin // n
            m.codes.data, m.codebooks.data, m.scales.data = weights(fin // n, fout, 1, 16, torch.float16)
            return m
        qkv = ShardedQuantizedLinearGroup([lin(8192, 8192), lin(8192, 1024), lin(8192, 1024)])
        o = lin(8192, 8192)
        gu = ShardedQuantizedLinearGroup([lin(8192, 28672), lin(8192, 28672)])
        down = lin(28672, 8192)
        x = torch.randn((1, 8192 // n), dtype=torch.float16, device=DEV)
        xi = torch.randn((1, 28672 // n), dtype=torch.float16, device=DEV)
        for _ in range(reps):
            qkv(x); o(x); gu(x); down(xi)
        torch.cuda.synchronize()
        dist.destroy_process_group()
        return
    else:
        raise SystemExit(f"unknown case {case}")
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
