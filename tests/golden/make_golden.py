"""Generates tests/golden/native_ops.npz by EXECUTING THE REFERENCE'S OWN OP FUNCTIONS.

Several hot-path rows of the reference are native torch calls wrapped in tiny functions at the
top of its benchmark scripts.  Those scripts cannot be imported (module-level sweeps that
require CUDA), so this script parses each file under /root/reference with `ast`, compiles only
the named FunctionDef, and calls it on seeded CPU inputs shaped like the script's own
(fp16 square matrices + int64 index, fp32 for sort / sparse).  Nothing is copied into the repo:
the reference is read at generation time only; the .npz holds inputs and outputs.

Functions that return None and keep their result in a local (op_native_scatter_add_
benchmark_scatter_add.py:22-25, op_native_scatter_multiply_ benchmark_scatter_multiply.py:42-45,
op_native_coalesce benchmark_sparse_coalesce.py:40-42) cannot yield a vector; for those the same
torch call the cited lines make is issued here directly (multiply accumulates into ones — the
torch_scatter.scatter_mul identity — because the script's zeros give an all-zero result).

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
torch version used is recorded in the file (the reference pins torch 1.11; these ops' CPU
semantics are unchanged).
"""
import ast
import os

import numpy as np
import torch

REF = "/root/reference/op_bm_scripts"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native_ops.npz")


def ref_function(script, name):
    """Compile one top-level function of a reference script (decorators dropped)."""
    path = os.path.join(REF, script)
    tree = ast.parse(open(path).read(), filename=path)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            node.decorator_list = []
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"torch": torch}
            exec(compile(mod, path, "exec"), ns)
            return ns[name]
    raise KeyError(f"{name} not found in {script}")


def main():
    torch.manual_seed(42)
    z = {"torch_version": np.array(torch.__version__)}

    # --- sort: op_native_sort(input, dim, stable)  benchmark_native_sort.py:28-30
    op_sort = ref_function("benchmark_native_sort.py", "op_native_sort")
    for tag, shape, dim in (("sort1d", (5000,), 0), ("sort2d_d0", (70, 90), 0), ("sort2d_d1", (70, 90), 1)):
        x = torch.rand(*shape)
        x = torch.nn.functional.dropout(x, p=0.5, training=True)  # the script's sparsity → many ties at 0
        v, i = op_sort(x, dim, True)
        z[tag + "_in"], z[tag + "_dim"], z[tag + "_val"], z[tag + "_idx"] = x.numpy(), np.array(dim), v.numpy(), i.numpy()

    # --- index_add_: op_native_index_add_(input, dim, index, source)  benchmark_native_index_add_.py:13-16 (dim=1, fp16)
    op_iadd = ref_function("benchmark_native_index_add_.py", "op_native_index_add_")
    L = 96
    inp = torch.rand(L, L, dtype=torch.float16)
    src = torch.rand(L, L, dtype=torch.float16)
    idx = torch.randint(0, L, (L,), dtype=torch.int64)
    out = inp.clone()
    op_iadd(out, 1, idx, src)
    scale = inp.float().clone().index_add_(1, idx, src.float())
    z["iadd_in"], z["iadd_src"], z["iadd_index"], z["iadd_out"], z["iadd_scale"] = \
        inp.numpy(), src.numpy(), idx.numpy(), out.numpy(), scale.numpy()

    # --- index_select: op_native_index_select(input, dim, index)  benchmark_native_index_select.py:12-15
    op_isel = ref_function("benchmark_native_index_select.py", "op_native_index_select")
    gelu = ref_function("benchmark_fused_index_select_reduce.py", "gelu")  # eager twin of the jit fn, :18-20
    inp = torch.rand(L, L, dtype=torch.float16)
    idx = torch.randint(0, L, (L // 2,), dtype=torch.int64)
    z["isel_in"], z["isel_index"] = inp.numpy(), idx.numpy()
    for dim in (0, 1):
        z[f"isel_out_d{dim}"] = op_isel(inp, dim, idx).numpy()
    z["isel_sum_d0"] = np.array(float(gelu(inp.float(), 0, idx)))
    z["isel_abs_sum_d0"] = np.array(float(gelu(inp.float().abs(), 0, idx)))

    # --- index_add → index_select → sum(dim): gelu(input, dim, index, other)  benchmark_fused_index_add_reduce.py:18-20
    gelu2 = ref_function("benchmark_fused_index_add_reduce.py", "gelu")
    a = torch.rand(L, L)
    b = torch.rand(L, L)
    idx = torch.randint(0, L, (L,), dtype=torch.int64)
    z["far_in"], z["far_other"], z["far_index"] = a.numpy(), b.numpy(), idx.numpy()
    for dim in (0, 1):
        z[f"far_out_d{dim}"] = gelu2(a, dim, idx, b).numpy()

    # --- spmm: op_native_smm(matA, matB)  benchmark_sparse_spmm.py:12-14 (COO fp32 × dense fp32)
    op_smm = ref_function("benchmark_sparse_spmm.py", "op_native_smm")
    m = n = 120
    A = torch.nn.functional.dropout(torch.rand(m, n), p=0.95, training=True).to_sparse()
    Bm = torch.rand(n, 40)
    z["smm_index"], z["smm_value"], z["smm_m"], z["smm_n"], z["smm_B"] = \
        A.indices().numpy(), A.values().numpy(), np.array(m), np.array(n), Bm.numpy()
    z["smm_out"] = op_smm(A, Bm).numpy()
    z["smm_scale"] = op_smm(torch.sparse_coo_tensor(A.indices(), A.values().abs(), (m, n)), Bm.abs()).numpy()

    # --- coalesce: mat.coalesce()  benchmark_sparse_coalesce.py:40-42; input built as :129-159 (RF=2)
    mat = torch.nn.functional.dropout(torch.rand(150, 130), p=0.9, training=True).to_sparse()
    index_, value_ = mat.indices(), mat.values()
    index = torch.cat((index_, index_), dim=1)
    value = torch.cat((value_, value_))
    index = index.index_select(1, torch.randperm(index.shape[1]))
    c = torch.sparse_coo_tensor(index, value, (150, 130)).coalesce()
    z["coal_index"], z["coal_value"], z["coal_m"], z["coal_n"] = index.numpy(), value.numpy(), np.array(150), np.array(130)
    z["coal_out_index"], z["coal_out_value"] = c.indices().numpy(), c.values().numpy()

    # --- native scatter_add_: temp.scatter_add_(dim, idx, src)  benchmark_scatter_add.py:22-25 (fp16, full-shape idx)
    s = torch.rand(L, L, dtype=torch.float16)
    i = torch.randint(0, L // 4, (L, L), dtype=torch.int64)
    z["sadd_src"], z["sadd_idx"] = s.numpy(), i.numpy()
    z["sadd_out"] = torch.zeros_like(s).scatter_add_(0, i, s).numpy()
    z["sadd_scale"] = torch.zeros(L, L).scatter_add_(0, i, s.float()).numpy()

    # --- native scatter_(reduce="multiply")  benchmark_scatter_multiply.py:42-45 (fp32, dim=-1), into ones
    s = torch.rand(40, 60) + 0.5
    i = torch.randint(0, 30, (40, 60), dtype=torch.int64)
    z["smul_src"], z["smul_idx"] = s.numpy(), i.numpy()
    z["smul_out"] = torch.ones_like(s).scatter_(-1, i, s, reduce="multiply").numpy()

    np.savez_compressed(OUT, **z)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
