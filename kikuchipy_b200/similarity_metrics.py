"""GPU similarity metrics with the reference's ``SimilarityMetric`` plugin surface.

Mirrors /root/reference/src/kikuchipy/indexing/similarity_metrics/:
``_similarity_metric.py:23-253`` (the ABC: constructor, properties, ``__repr__``,
``raise_error_if_invalid``), ``_normalized_cross_correlation.py:26-183`` and
``_normalized_dot_product.py:25-174`` (``__call__``, ``prepare_experimental``,
``prepare_dictionary``, ``match``).

The classes subclass the *real* ``kikuchipy.indexing.SimilarityMetric`` when kikuchipy is
importable, so an unmodified ``EBSD.dictionary_indexing(metric=...)`` accepts them
(``signals/ebsd.py:3067`` does an ``isinstance`` check); otherwise they subclass the replica
below.  ``prepare_*`` return device-resident handles, ``match`` returns a lazy block whose
``topk`` / ``argtopk`` run the fused tensor-core GEMM + selection + exact rescoring and whose
``compute()`` / ``np.asarray`` give the exact float32 block.  All arithmetic is CUDA; there is
no CPU fallback.
"""

from __future__ import annotations

import abc

import numpy as np

from . import _lib

try:  # pragma: no cover - kikuchipy is not installed in the build container
    from kikuchipy.indexing import SimilarityMetric as _ReferenceABC
except Exception:  # noqa: BLE001
    _ReferenceABC = None


class _SimilarityMetricReplica(abc.ABC):
    """Replica of ``kikuchipy.indexing.SimilarityMetric`` (same attributes and behaviour)."""

    _allowed_dtypes: list = []
    _sign = None

    def __init__(
        self,
        n_experimental_patterns=None,
        n_dictionary_patterns=None,
        navigation_mask=None,
        signal_mask=None,
        dtype="float32",
        rechunk=False,
    ):
        self._n_experimental_patterns = n_experimental_patterns
        self._n_dictionary_patterns = n_dictionary_patterns
        self._navigation_mask = navigation_mask
        self._signal_mask = signal_mask
        self._dtype = np.dtype(dtype)
        self._rechunk = rechunk

    def __repr__(self):
        string = f"{self.__class__.__name__}: {np.dtype(self.dtype).name}, "
        sign_string = {1: "greater is better", -1: "lower is better"}
        string += sign_string[self.sign]
        string += f", rechunk: {self.rechunk}, "
        string += f"navigation mask: {self.navigation_mask is not None}, "
        string += f"signal mask: {self.signal_mask is not None}"
        return string

    @property
    def allowed_dtypes(self):
        return self._allowed_dtypes

    @property
    def dtype(self):
        return self._dtype

    @dtype.setter
    def dtype(self, value):
        self._dtype = np.dtype(value)

    @property
    def n_dictionary_patterns(self):
        return self._n_dictionary_patterns

    @n_dictionary_patterns.setter
    def n_dictionary_patterns(self, value):
        self._n_dictionary_patterns = value

    @property
    def n_experimental_patterns(self):
        return self._n_experimental_patterns

    @n_experimental_patterns.setter
    def n_experimental_patterns(self, value):
        self._n_experimental_patterns = value

    @property
    def navigation_mask(self):
        return self._navigation_mask

    @navigation_mask.setter
    def navigation_mask(self, value):
        self._navigation_mask = value

    @property
    def signal_mask(self):
        return self._signal_mask

    @signal_mask.setter
    def signal_mask(self, value):
        self._signal_mask = value

    @property
    def sign(self):
        return self._sign

    @property
    def rechunk(self):
        return self._rechunk

    @rechunk.setter
    def rechunk(self, value):
        self._rechunk = value

    @abc.abstractmethod
    def prepare_dictionary(self, *args, **kwargs):
        return NotImplemented  # pragma: no cover

    @abc.abstractmethod
    def prepare_experimental(self, *args, **kwargs):
        return NotImplemented  # pragma: no cover

    @abc.abstractmethod
    def match(self, *args, **kwargs):
        return NotImplemented  # pragma: no cover

    def raise_error_if_invalid(self):
        allowed_dtypes = self.allowed_dtypes
        if len(allowed_dtypes) != 0 and self.dtype not in allowed_dtypes:
            raise ValueError(
                f"Data type {self.dtype} not among supported data types {allowed_dtypes}"
            )


SimilarityMetric = _ReferenceABC if _ReferenceABC is not None else _SimilarityMetricReplica


class SimilarityBlock:
    """Lazy ``(n_experimental, n_dictionary)`` similarity block returned by ``match()``.

    Offers what the reference driver uses on the Dask array it gets from ``match``
    (``_dictionary_indexing.py:197-201``): ``argtopk`` / ``topk`` along the last axis, whose
    results are NumPy arrays (``reshape`` works; ``dask.compute`` passes them through).
    """

    def __init__(self, ctx, experimental, dictionary):
        self._ctx = ctx
        self._exp = experimental
        self._dict = dictionary
        self.shape = (experimental.shape[0], dictionary.shape[0])
        self.dtype = np.dtype(np.float32)
        self._cache = {}

    def _topk(self, k):
        k = int(k)
        if k < 0:
            raise NotImplementedError("only the k largest values are supported (k > 0)")
        if k not in self._cache:
            self._cache = {k: self._ctx.match_topk(self._exp, self._dict, k)}
        return self._cache[k]

    def topk(self, k, axis=-1):
        if axis not in (-1, 1):
            raise ValueError("top-k is taken along the dictionary axis (axis=-1)")
        return self._topk(k)[1]

    def argtopk(self, k, axis=-1):
        if axis not in (-1, 1):
            raise ValueError("top-k is taken along the dictionary axis (axis=-1)")
        return self._topk(k)[0]

    def compute(self, **kwargs):
        return self._ctx.match_full(self._exp, self._dict)

    def __array__(self, dtype=None, copy=None):
        out = self.compute()
        return out if dtype is None else out.astype(dtype)


class _Prepared64:
    """What ``prepare_*`` returns in float64 mode: the float32 prepared set (candidate nomination) and the
    raw rows on the device (final float64 scores)."""

    def __init__(self, patterns32, raw, row_index):
        self.patterns32 = patterns32
        self.raw = raw              # CUDA tensor (source rows, -1)
        self.row_index = row_index  # source row of every kept row, or None
        self.shape = patterns32.shape

    def compute(self):
        return self


class SimilarityBlock64:
    """``match()`` of two float64-mode sets.  ``topk`` / ``argtopk``: the float32 pipeline nominates
    ``k + 16`` candidates per row (tensor-core candidates, exact float32 rescoring, certificate), the
    float64 kernel gives them the reference's float64 scores and the order follows those.  A row whose
    k-th float64 score does not clear the best score that can hide outside the nominated set by the
    float32 rounding level is done again with twice as many candidates (exact duplicates in the
    dictionary end at k' = N, i.e. every pair in float64)."""

    MARGIN = 4e-6  # bound on |float32 exact score - float64 score| (unit vectors, fp32 FMA chain + rounding of the rows)

    def __init__(self, ctx, experimental, dictionary, kdi_metric):
        self._ctx, self._exp, self._dict, self._metric = ctx, experimental, dictionary, kdi_metric
        self.shape = (experimental.shape[0], dictionary.shape[0])
        self.dtype = np.dtype(np.float64)
        self._cache = {}

    def _topk(self, k):
        k = int(k)
        if k < 0:
            raise NotImplementedError("only the k largest values are supported (k > 0)")
        if k in self._cache:
            return self._cache[k]
        m, n = self.shape
        extra = 16
        while True:
            kk = min(n, k + extra)
            idx32, sc32 = self._ctx.match_topk(self._exp.patterns32, self._dict.patterns32, kk)
            sc64 = self._ctx.scores_f64(self._exp.raw, self._exp.row_index, self._dict.raw, self._metric, idx32)
            # rank by (float64 score descending, index ascending); NaN rows keep the float32 order
            key = np.where(np.isnan(sc64), -np.inf, sc64)
            order = np.lexsort((idx32, -key), axis=1)
            idx = np.take_along_axis(idx32, order, axis=1)[:, :k]
            sc = np.take_along_axis(sc64, order, axis=1)[:, :k]
            if kk == n:
                break
            ok = np.isnan(sc[:, -1]) | np.isnan(sc32[:, -1]) | (sc[:, -1] > sc32[:, -1].astype(np.float64) + self.MARGIN)
            if ok.all():
                break
            extra = 2 * (kk - k) + 16
        self._cache = {k: (idx, sc)}
        return self._cache[k]

    def topk(self, k, axis=-1):
        if axis not in (-1, 1):
            raise ValueError("top-k is taken along the dictionary axis (axis=-1)")
        return self._topk(k)[1]

    def argtopk(self, k, axis=-1):
        if axis not in (-1, 1):
            raise ValueError("top-k is taken along the dictionary axis (axis=-1)")
        return self._topk(k)[0]

    def compute(self, **kwargs):
        m, n = self.shape
        cand = np.broadcast_to(np.arange(n, dtype=np.int64), (m, n))
        return self._ctx.scores_f64(self._exp.raw, self._exp.row_index, self._dict.raw, self._metric, cand)

    def __array__(self, dtype=None, copy=None):
        out = self.compute()
        return out if dtype is None else out.astype(dtype)


class _GpuMetric(SimilarityMetric):
    # float64: candidates from the float32 pipeline, final scores and order in float64 (SimilarityBlock64)
    _allowed_dtypes = [np.float32, np.float64]
    _sign = 1
    _kdi_metric = None

    def __init__(self, *args, context=None, **kwargs):
        super().__init__(*args, **kwargs)
        self._context = context

    # -- device plumbing ------------------------------------------------------
    @property
    def context(self):
        if self._context is None:
            self._context = _lib.default_context()
        return self._context

    def _sync_signal_mask(self):
        self.context.set_signal_mask(self.signal_mask)

    def __call__(self, experimental, dictionary):
        """Similarities between experimental and dictionary patterns as a NumPy array
        (``_normalized_cross_correlation.py:64-86``)."""
        experimental = self.prepare_experimental(experimental)
        dictionary = np.asarray(dictionary) if not hasattr(dictionary, "data_ptr") else dictionary
        dictionary = dictionary.reshape((self.n_dictionary_patterns, -1))
        dictionary = self.prepare_dictionary(dictionary)
        return self.match(experimental, dictionary).compute()

    def prepare_experimental(self, patterns):
        """cast -> reshape ``(n_experimental_patterns, -1)`` -> drop navigation-masked rows ->
        signal mask -> normalise (``_normalized_cross_correlation.py:88-128``)."""
        self.raise_error_if_invalid()
        self._sync_signal_mask()
        n = int(self.n_experimental_patterns)
        if np.dtype(self.dtype) == np.float64:
            raw = self.context.device_rows(patterns, n)
            keep = None
            if self.navigation_mask is not None:
                keep = np.flatnonzero(~np.asarray(self.navigation_mask, dtype=bool).ravel())
            return _Prepared64(self.context.patterns(raw, n, self._kdi_metric, self.navigation_mask), raw, keep)
        return self.context.patterns(patterns, n, self._kdi_metric, self.navigation_mask)

    def prepare_dictionary(self, patterns):
        """cast -> signal mask -> normalise; ``patterns`` is 2-D
        (``_normalized_cross_correlation.py:130-159``)."""
        self.raise_error_if_invalid()
        self._sync_signal_mask()
        if len(patterns.shape) != 2:
            raise ValueError("dictionary patterns must be reshaped to (n patterns, n pixels) first")
        if np.dtype(self.dtype) == np.float64:
            raw = self.context.device_rows(patterns, int(patterns.shape[0]))
            return _Prepared64(self.context.patterns(raw, int(patterns.shape[0]), self._kdi_metric, None), raw, None)
        return self.context.patterns(patterns, int(patterns.shape[0]), self._kdi_metric, None)

    def match(self, experimental, dictionary):
        """``einsum("ik,mk->im")`` of prepared sets (``_normalized_cross_correlation.py:161-183``),
        evaluated lazily."""
        if isinstance(experimental, _Prepared64) != isinstance(dictionary, _Prepared64):
            raise ValueError("experimental and dictionary patterns were prepared with different metric dtypes")
        if isinstance(experimental, _Prepared64):
            self._sync_signal_mask()
            return SimilarityBlock64(self.context, experimental, dictionary, self._kdi_metric)
        return SimilarityBlock(self.context, experimental, dictionary)


class NormalizedCrossCorrelationMetric(_GpuMetric):
    r"""Normalized cross-correlation (Pearson) on the GPU:
    :math:`r = \sum (x_i-\bar x)(y_i-\bar y) / (\|x-\bar x\| \|y-\bar y\|)`.

    Drop-in for ``kikuchipy.indexing.NormalizedCrossCorrelationMetric``
    (``_normalized_cross_correlation.py:26``).  ``dtype=float64`` like the reference: candidates are
    nominated by the float32 pipeline, their final scores and order are computed in float64 on the
    device (``SimilarityBlock64``)."""

    _kdi_metric = _lib.KDI_NCC


class NormalizedDotProductMetric(_GpuMetric):
    r"""Normalized dot product on the GPU: :math:`\rho = \langle x, y\rangle / (\|x\| \|y\|)`,
    no centring (``_normalized_dot_product.py:25,181-194``)."""

    _kdi_metric = _lib.KDI_NDP
