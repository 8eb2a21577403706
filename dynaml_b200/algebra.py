"""Host-side mirror of the blocked containers the GP path exchanges with its callers.

Reference (CORE/algebra): PartitionedMatrix.scala:19-287 (PartitionedMatrix, LowerTriPartitionedMatrix,
PartitionedPSDMatrix), PartitionedVector.scala:21-158.  The reference keeps a lazy Stream of ((i, j), DenseMatrix)
blocks; here a block is a column-major numpy array filled by the device (dgp_data_kernel_block) and the stream is a
list.  Only what the kernel-matrix builders and their callers touch is mirrored: dimensions, block access, `_data`,
range slicing re-indexed from 0 (PartitionedMatrix.scala:50-55), transposition and the dense view (`toBreezeMatrix`).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

Block = Tuple[Tuple[int, int], np.ndarray]


class PartitionedMatrix:
    def __init__(self, data: Iterable[Block], num_rows: int = -1, num_cols: int = -1, num_row_blocks: int = -1,
                 num_col_blocks: int = -1):
        self._underlying: List[Block] = sorted(((int(i), int(j)), np.asarray(b)) for (i, j), b in data)
        keys = [k for k, _ in self._underlying]
        self.rowBlocks = num_row_blocks if num_row_blocks >= 0 else 1 + max(i for i, _ in keys)
        self.colBlocks = num_col_blocks if num_col_blocks >= 0 else 1 + max(j for _, j in keys)
        blocks = self._blocks()
        self.rows = num_rows if num_rows >= 0 else sum(blocks[(i, 0)].shape[0] for i in range(self.rowBlocks))
        self.cols = num_cols if num_cols >= 0 else sum(blocks[(0, j)].shape[1] for j in range(self.colBlocks))

    # ---- the reference's accessors -----------------------------------------------------------------------------
    @property
    def _data(self) -> List[Block]:
        return list(self._underlying)

    def _blocks(self) -> Dict[Tuple[int, int], np.ndarray]:
        return dict(self._data)

    def __call__(self, rows: range, cols: range) -> "PartitionedMatrix":
        """Blocks r in `rows`, c in `cols`, re-indexed from 0 (PartitionedMatrix.scala:50-55)."""
        r0, c0 = rows.start, cols.start
        sel = [((i - r0, j - c0), b) for (i, j), b in self._data if i in rows and j in cols]
        return PartitionedMatrix(sel, num_row_blocks=len(rows), num_col_blocks=len(cols))

    @property
    def t(self) -> "PartitionedMatrix":
        return PartitionedMatrix((((j, i), b.T) for (i, j), b in self._data), self.cols, self.rows, self.colBlocks,
                                 self.rowBlocks)

    def toBreezeMatrix(self) -> np.ndarray:
        """Dense column-major copy (PartitionedMatrix.scala:91-94)."""
        blocks = self._blocks()
        return np.asfortranarray(np.block([[blocks[(i, j)] for j in range(self.colBlocks)] for i in range(self.rowBlocks)]))

    toDenseMatrix = toBreezeMatrix

    def filterBlocks(self, f) -> List[Block]:
        return [(k, b) for k, b in self._data if f(k)]


class LowerTriPartitionedMatrix(PartitionedMatrix):
    """Lower block triangle; the blocks above the diagonal are zeros of the right shape (PartitionedMatrix.scala:195-246)."""

    def __init__(self, underlying: Iterable[Block], num_rows: int = -1, num_cols: int = -1, num_row_blocks: int = -1,
                 num_col_blocks: int = -1):
        lower = {(int(i), int(j)): np.asarray(b) for (i, j), b in underlying}
        self._lower = lower
        nb = 1 + max(i for i, _ in lower)
        full = dict(lower)
        for i in range(nb):
            for j in range(i + 1, nb):
                full[(i, j)] = np.zeros((lower[(i, i)].shape[0], lower[(j, j)].shape[1]), order="F")
        super().__init__(full.items(), num_rows, num_cols, num_row_blocks, num_col_blocks)

    @property
    def _underlyingdata(self) -> List[Block]:
        return sorted(self._lower.items())


class PartitionedPSDMatrix(PartitionedMatrix):
    """Symmetric matrix stored as its lower block triangle; the upper blocks are the transposes
    (PartitionedMatrix.scala:248-259).  `_underlyingdata` is what the builder produced."""

    def __init__(self, underlying: Iterable[Block], num_rows: int = -1, num_cols: int = -1, num_row_blocks: int = -1,
                 num_col_blocks: int = -1):
        lower = {(int(i), int(j)): np.asarray(b) for (i, j), b in underlying}
        assert all(i >= j for i, j in lower), "a PartitionedPSDMatrix is built from blocks on or below the diagonal"
        self._lower = lower
        full = dict(lower)
        for (i, j), b in lower.items():
            if i > j:
                full[(j, i)] = b.T
        super().__init__(full.items(), num_rows, num_cols, num_row_blocks, num_col_blocks)

    @property
    def _underlyingdata(self) -> List[Block]:
        return sorted(self._lower.items())


class PartitionedVector:
    """PartitionedVector.scala:21-158: a column vector as a list of (block index, segment)."""

    def __init__(self, data: Iterable[Tuple[int, np.ndarray]], num_rows: int = -1, num_row_blocks: int = -1):
        self._underlying = sorted((int(i), np.asarray(v, dtype=np.float64).ravel()) for i, v in data)
        self.rowBlocks = num_row_blocks if num_row_blocks >= 0 else len(self._underlying)
        self.rows = num_rows if num_rows >= 0 else sum(v.size for _, v in self._underlying)

    @classmethod
    def fromDense(cls, v: Sequence[float], num_elements_per_block: int) -> "PartitionedVector":
        v = np.asarray(v, dtype=np.float64).ravel()
        segs = [(b, v[s:s + num_elements_per_block]) for b, s in enumerate(range(0, v.size, num_elements_per_block))]
        return cls(segs, v.size, len(segs))

    @property
    def _data(self) -> List[Tuple[int, np.ndarray]]:
        return list(self._underlying)

    def __call__(self, r: range) -> "PartitionedVector":
        return PartitionedVector([(i - r.start, v) for i, v in self._data if i in r], num_row_blocks=len(r))

    def toBreezeVector(self) -> np.ndarray:
        return np.concatenate([v for _, v in self._data]) if self._underlying else np.zeros(0)

    toDenseVector = toBreezeVector
