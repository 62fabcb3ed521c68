"""Known-answer vectors of the upstream packages' OWN test suites (pytorch_scatter
test/test_scatter.py and test/test_segment.py — the `tests = [...]` tables every upstream reduce is
checked against).  The packages are absent from this container and there is no network, so the
tables are restated here from the upstream repositories as the author knows them; they were NOT
fetched, and every expected value below is small enough to verify by hand from the documented
semantics (the derivation is in the comments), so each stands as a known-answer test on its own.

What they pin beyond the README examples (tests/golden/upstream_published.py): sum / mul / mean of
every shape class (1-D, 1-D index over a 2-D src along dim 0, full-shape 2-D index along dim 1, an
index with FEWER dims than src that upstream broadcasts over the trailing dim), mul's identity 1
and mean's / min's / max's 0 in an empty bucket, the arg sentinel = src.size(dim), arg = position
along `dim`, and segment_coo / segment_csr / gather_coo / gather_csr (sorted index / row pointers, per-row
pointers, an empty segment in the middle and at the end; test/test_gather.py).  Upstream runs them with integer and floating dtypes; the
library's scatter covers floating values, so they are used as float32 here."""
import torch

T = torch.tensor

# ---- test/test_scatter.py -----------------------------------------------------------------------
SCATTER = [
    # bucket 0 <- {1@0, 2@2}, bucket 1 <- {3@1, 4@3, 5@4}, bucket 2 empty, bucket 3 <- {6@5}
    dict(src=T([1., 3., 2., 4., 5., 6.]), index=T([0, 1, 0, 1, 1, 3]), dim=-1,
         sum=T([3., 12., 0., 6.]), mul=T([2., 60., 1., 6.]), mean=T([1.5, 4., 0., 6.]),
         min=T([1., 3., 0., 6.]), arg_min=T([0, 1, 6, 5]),
         max=T([2., 5., 0., 6.]), arg_max=T([2, 4, 6, 5])),
    # the same buckets over rows of two columns (1-D index along dim 0)
    dict(src=T([[1., 2.], [5., 6.], [3., 4.], [7., 8.], [9., 10.], [11., 12.]]), index=T([0, 1, 0, 1, 1, 3]), dim=0,
         sum=T([[4., 6.], [21., 24.], [0., 0.], [11., 12.]]),
         mul=T([[3., 8.], [315., 480.], [1., 1.], [11., 12.]]),
         mean=T([[2., 3.], [7., 8.], [0., 0.], [11., 12.]]),
         min=T([[1., 2.], [5., 6.], [0., 0.], [11., 12.]]), arg_min=T([[0, 0], [1, 1], [6, 6], [5, 5]]),
         max=T([[3., 4.], [9., 10.], [0., 0.], [11., 12.]]), arg_max=T([[2, 2], [4, 4], [6, 6], [5, 5]])),
    # full-shape 2-D index along dim 1.  row 1: bucket 0 <- {2@0, 4@1, 6@3}, bucket 1 <- {8@2, 10@4},
    # bucket 2 <- {12@5}, bucket 3 empty
    dict(src=T([[1., 5., 3., 7., 9., 11.], [2., 4., 8., 6., 10., 12.]]),
         index=T([[0, 1, 0, 1, 1, 3], [0, 0, 1, 0, 1, 2]]), dim=1,
         sum=T([[4., 21., 0., 11.], [12., 18., 12., 0.]]),
         mul=T([[3., 315., 1., 11.], [48., 80., 12., 1.]]),
         mean=T([[2., 7., 0., 11.], [4., 9., 12., 0.]]),
         min=T([[1., 5., 0., 11.], [2., 8., 12., 0.]]), arg_min=T([[0, 1, 6, 5], [0, 2, 5, 6]]),
         max=T([[3., 9., 0., 11.], [6., 10., 12., 0.]]), arg_max=T([[2, 4, 6, 5], [3, 4, 5, 6]])),
    # index [2, 3] against src [2, 3, 2]: upstream broadcasts it over the trailing dim; 3 buckets
    # (max index 2), sentinel = src.size(1) = 3.  batch 1: bucket 0 <- row 1, bucket 2 <- rows 0, 2
    dict(src=T([[[1., 2.], [5., 6.], [3., 4.]], [[10., 11.], [7., 9.], [12., 13.]]]),
         index=T([[0, 1, 0], [2, 0, 2]]), dim=1,
         sum=T([[[4., 6.], [5., 6.], [0., 0.]], [[7., 9.], [0., 0.], [22., 24.]]]),
         mul=T([[[3., 8.], [5., 6.], [1., 1.]], [[7., 9.], [1., 1.], [120., 143.]]]),
         mean=T([[[2., 3.], [5., 6.], [0., 0.]], [[7., 9.], [0., 0.], [11., 12.]]]),
         min=T([[[1., 2.], [5., 6.], [0., 0.]], [[7., 9.], [0., 0.], [10., 11.]]]),
         arg_min=T([[[0, 0], [1, 1], [3, 3]], [[1, 1], [3, 3], [0, 0]]]),
         max=T([[[3., 4.], [5., 6.], [0., 0.]], [[7., 9.], [0., 0.], [12., 13.]]]),
         arg_max=T([[[2, 2], [1, 1], [3, 3]], [[1, 1], [3, 3], [2, 2]]])),
    # everything into one bucket
    dict(src=T([[1., 3.], [2., 4.]]), index=T([[0, 0], [0, 0]]), dim=1,
         sum=T([[4.], [6.]]), mul=T([[3.], [8.]]), mean=T([[2.], [3.]]),
         min=T([[1.], [2.]]), arg_min=T([[0], [0]]), max=T([[3.], [4.]]), arg_max=T([[1], [1]])),
]

# ---- test/test_segment.py (segment_coo over `index`, segment_csr over `indptr`) ------------------
SEGMENT = [
    # segments [0,2) = {1,2}, [2,5) = {3,4,5}, [5,5) empty, [5,6) = {6}
    dict(src=T([1., 2., 3., 4., 5., 6.]), index=T([0, 0, 1, 1, 1, 3]), indptr=T([0, 2, 5, 5, 6]),
         sum=T([3., 12., 0., 6.]), mean=T([1.5, 4., 0., 6.]),
         min=T([1., 3., 0., 6.]), arg_min=T([0, 2, 6, 5]), max=T([2., 5., 0., 6.]), arg_max=T([1, 4, 6, 5])),
    dict(src=T([[1., 2.], [3., 4.], [5., 6.], [7., 8.], [9., 10.], [11., 12.]]), index=T([0, 0, 1, 1, 1, 3]),
         indptr=T([0, 2, 5, 5, 6]),
         sum=T([[4., 6.], [21., 24.], [0., 0.], [11., 12.]]), mean=T([[2., 3.], [7., 8.], [0., 0.], [11., 12.]]),
         min=T([[1., 2.], [5., 6.], [0., 0.], [11., 12.]]), arg_min=T([[0, 0], [2, 2], [6, 6], [5, 5]]),
         max=T([[3., 4.], [9., 10.], [0., 0.], [11., 12.]]), arg_max=T([[1, 1], [4, 4], [6, 6], [5, 5]])),
    # per-row pointers.  row 1: [0,3) = {2,4,6}, [3,5) = {8,10}, [5,6) = {12}, [6,6) empty
    dict(src=T([[1., 3., 5., 7., 9., 11.], [2., 4., 6., 8., 10., 12.]]),
         index=T([[0, 0, 1, 1, 1, 3], [0, 0, 0, 1, 1, 2]]), indptr=T([[0, 2, 5, 5, 6], [0, 3, 5, 6, 6]]),
         sum=T([[4., 21., 0., 11.], [12., 18., 12., 0.]]), mean=T([[2., 7., 0., 11.], [4., 9., 12., 0.]]),
         min=T([[1., 5., 0., 11.], [2., 8., 12., 0.]]), arg_min=T([[0, 2, 6, 5], [0, 3, 5, 6]]),
         max=T([[3., 9., 0., 11.], [6., 10., 12., 0.]]), arg_max=T([[1, 4, 6, 5], [2, 4, 5, 6]])),
]

# ---- test/test_gather.py (gather_coo over `index`, gather_csr over `indptr`) ---------------------
# expected[..., e] = src[..., segment of e]; segment 2 is empty and segment 3 holds the last element
GATHER = [
    dict(src=T([1., 2., 3., 4.]), index=T([0, 0, 1, 1, 1, 3]), indptr=T([0, 2, 5, 5, 6]),
         expected=T([1., 1., 2., 2., 2., 4.])),
    dict(src=T([[1., 2.], [3., 4.], [5., 6.], [7., 8.]]), index=T([0, 0, 1, 1, 1, 3]), indptr=T([0, 2, 5, 5, 6]),
         expected=T([[1., 2.], [1., 2.], [3., 4.], [3., 4.], [3., 4.], [7., 8.]])),
    dict(src=T([[1., 3., 5., 7.], [2., 4., 6., 8.]]), index=T([[0, 0, 1, 1, 1, 3], [0, 0, 0, 1, 1, 2]]),
         indptr=T([[0, 2, 5, 5, 6], [0, 3, 5, 6, 6]]),
         expected=T([[1., 1., 3., 3., 3., 7.], [2., 2., 2., 4., 4., 6.]])),
]
