"""Known-answer vectors PUBLISHED by the upstream packages the reference pins (requirements.txt:212-213):
the worked examples of the pytorch_scatter README (scatter_max) and the pytorch_sparse README
(coalesce, transpose, spmm).  The packages themselves are absent from this container, so the
vectors are restated from those READMEs and each was re-derived by hand from the documented
semantics (see the comments); they pin the points the oracle otherwise only restates: the arg
sentinel (= src.size(dim)) and zero fill of empty buckets, a genuine 0 maximum keeping its real
arg, (row, col) ordering and duplicate summation in coalesce / transpose, and spmm.

(Upstream's scatter_max example uses an integer src; the library covers floating dtypes, so the
same numbers are used as float32.)"""
import torch

# pytorch_scatter README: out, argmax = scatter_max(src, index, dim=-1)
SCATTER_MAX = dict(
    src=torch.tensor([[2., 0., 1., 4., 3.], [0., 2., 1., 3., 4.]]),
    index=torch.tensor([[4, 5, 4, 2, 3], [0, 0, 2, 2, 1]]),
    # row 0: bucket 4 <- {2@0, 1@2}, bucket 5 <- {0@1}, bucket 2 <- {4@3}, bucket 3 <- {3@4}; 0, 1 empty
    out=torch.tensor([[0., 0., 4., 3., 2., 0.], [2., 4., 3., 0., 0., 0.]]),
    arg=torch.tensor([[5, 5, 3, 4, 0, 1], [1, 4, 3, 5, 5, 5]]),
)

# pytorch_sparse README: coalesce(index, value, m=3, n=2)
_INDEX = torch.tensor([[1, 0, 1, 0, 2, 1], [0, 1, 1, 1, 0, 0]])
_VALUE = torch.tensor([[1., 2.], [2., 3.], [3., 4.], [4., 5.], [5., 6.], [6., 7.]])
COALESCE = dict(
    index=_INDEX, value=_VALUE, m=3, n=2,
    # (0,1) = [2,3]+[4,5]; (1,0) = [1,2]+[6,7]; (1,1); (2,0)
    out_index=torch.tensor([[0, 1, 1, 2], [1, 0, 1, 0]]),
    out_value=torch.tensor([[6., 8.], [7., 9.], [3., 4.], [5., 6.]]),
)

# pytorch_sparse README: transpose(index, value, 3, 2) of the same (uncoalesced) matrix
TRANSPOSE = dict(
    index=_INDEX, value=_VALUE, m=3, n=2,
    out_index=torch.tensor([[0, 0, 1, 1], [1, 2, 0, 1]]),
    out_value=torch.tensor([[7., 9.], [5., 6.], [6., 8.], [3., 4.]]),
)

# pytorch_sparse README: spmm(index, value, 3, 3, matrix)
SPMM = dict(
    index=torch.tensor([[0, 0, 1, 2, 2], [0, 2, 1, 0, 1]]),
    value=torch.tensor([1., 2., 4., 1., 3.]),
    m=3, n=3,
    matrix=torch.tensor([[1., 4.], [2., 5.], [3., 6.]]),
    out=torch.tensor([[7., 16.], [8., 20.], [7., 19.]]),
)
