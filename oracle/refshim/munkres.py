"""TEST INFRASTRUCTURE ONLY -- stand-in for the third-party ``munkres`` package.

The reference pins ``munkres==1.1.4`` (/root/reference/pyproject.toml:20,
poetry.lock:1359-1365) and calls it from ``py_max_match``
(/root/reference/src/keypoints/grouping.py:55-59,130).  The package is not
vendored in the reference and cannot be installed here (no network), so this
module restates the *published* algorithm of ``Munkres.compute`` from memory of
the 1.1.x sources.  PARITY UNPINNED: no genuine munkres wheel was available to
diff against; the optimal cost is cross-checked against
``scipy.optimize.linear_sum_assignment`` in tests/test_oracle.py
(``test_munkres_*``).

Only what the reference uses is provided: ``Munkres().compute(matrix)`` on a
rows<=cols float64 numpy matrix.  The two places whose exact shape decides the
tie-breaking of equal-cost assignments are kept as single functions
(``_find_a_zero`` and ``_step6``) so they can be re-pinned if a real 1.1.4
source ever becomes available.  The same control flow is restated in C++
(oracle/hpd_oracle.cpp: munkres_compute) and in CUDA (csrc/group.cu).

This directory is put on ``sys.path`` only by oracle/gen_golden.py and by the
tests that import the unmodified reference ``grouping.py``; product code never
imports it.
"""
import sys

__version__ = "1.1.4-restated"


class Munkres:
    def pad_matrix(self, matrix, pad_value=0):
        # Square the matrix with rows of ``pad_value``.  Real rows stay the
        # caller's objects (numpy row views for the reference) when they are
        # already wide enough, which is always the case for the reference
        # (grouping.py:126-128 pads columns first so rows <= cols).
        width = max(len(r) for r in matrix)
        n = max(width, len(matrix))
        out = []
        for r in matrix:
            r2 = r[:]
            if n > len(r):
                r2 = list(r2) + [pad_value] * (n - len(r))
            out.append(r2)
        while len(out) < n:
            out.append([pad_value] * n)
        return out

    def compute(self, cost_matrix):
        self.C = self.pad_matrix(cost_matrix)
        self.n = n = len(self.C)
        self.original_length = len(cost_matrix)
        self.original_width = len(cost_matrix[0])
        self.row_covered = [False] * n
        self.col_covered = [False] * n
        self.Z0_r = self.Z0_c = 0
        self.path = [[0, 0] for _ in range(2 * n)]
        self.marked = [[0] * n for _ in range(n)]

        table = {1: self._step1, 2: self._step2, 3: self._step3,
                 4: self._step4, 5: self._step5, 6: self._step6}
        step = 1
        while step in table:
            step = table[step]()

        return [(i, j)
                for i in range(self.original_length)
                for j in range(self.original_width)
                if self.marked[i][j] == 1]

    # -- step 1: subtract each row's minimum ---------------------------------
    def _step1(self):
        C, n = self.C, self.n
        for i in range(n):
            m = min(C[i])
            for j in range(n):
                C[i][j] -= m
        return 2

    # -- step 2: greedy initial stars, rows ascending, first free zero ---------
    def _step2(self):
        C, n = self.C, self.n
        for i in range(n):
            for j in range(n):
                if C[i][j] == 0 and not self.col_covered[j] and not self.row_covered[i]:
                    self.marked[i][j] = 1
                    self.col_covered[j] = True
                    self.row_covered[i] = True
                    break
        self._clear_covers()
        return 3

    # -- step 3: cover starred columns; all covered -> done ---------------------
    def _step3(self):
        n = self.n
        count = 0
        for i in range(n):
            for j in range(n):
                if self.marked[i][j] == 1 and not self.col_covered[j]:
                    self.col_covered[j] = True
                    count += 1
        return 7 if count >= n else 4

    # -- step 4: prime uncovered zeros until one has no star in its row ---------
    def _step4(self):
        row = col = 0
        while True:
            row, col = self._find_a_zero(row, col)
            if row < 0:
                return 6
            self.marked[row][col] = 2
            star_col = self._find_in_row(row, 1)
            if star_col >= 0:
                col = star_col
                self.row_covered[row] = True
                self.col_covered[col] = False
            else:
                self.Z0_r, self.Z0_c = row, col
                return 5

    # -- step 5: augment along the alternating path starting at Z0 --------------
    def _step5(self):
        path = self.path
        count = 0
        path[0][0], path[0][1] = self.Z0_r, self.Z0_c
        while True:
            r = self._find_in_col(path[count][1], 1)
            if r < 0:
                break
            count += 1
            path[count][0], path[count][1] = r, path[count - 1][1]
            c = self._find_in_row(path[count][0], 2)
            count += 1
            path[count][0], path[count][1] = path[count - 1][0], c
        for k in range(count + 1):
            r, c = path[k]
            self.marked[r][c] = 0 if self.marked[r][c] == 1 else 1
        self._clear_covers()
        for i in range(self.n):
            for j in range(self.n):
                if self.marked[i][j] == 2:
                    self.marked[i][j] = 0
        return 3

    # -- step 6: shift by the smallest uncovered value --------------------------
    def _step6(self):
        C, n = self.C, self.n
        m = self._find_smallest()
        for i in range(n):
            for j in range(n):
                # order matters bitwise: a covered-row / uncovered-col cell
                # becomes (C + m) - m, which is not always C again.
                if self.row_covered[i]:
                    C[i][j] += m
                if not self.col_covered[j]:
                    C[i][j] -= m
        return 4

    def _find_smallest(self):
        m = sys.maxsize
        for i in range(self.n):
            if self.row_covered[i]:
                continue
            for j in range(self.n):
                if not self.col_covered[j] and m > self.C[i][j]:
                    m = self.C[i][j]
        return m

    def _find_a_zero(self, i0=0, j0=0):
        # Rows cyclically from i0; inside a row columns cyclically from j0 with
        # NO early exit: the last uncovered zero of that cyclic order wins.
        # The search stops after the first row that produced a hit.
        n = self.n
        row = col = -1
        i = i0
        done = False
        while not done:
            j = j0
            while True:
                if self.C[i][j] == 0 and not self.row_covered[i] and not self.col_covered[j]:
                    row, col, done = i, j, True
                j = (j + 1) % n
                if j == j0:
                    break
            i = (i + 1) % n
            if i == i0:
                done = True
        return row, col

    def _find_in_row(self, row, mark):
        for j in range(self.n):
            if self.marked[row][j] == mark:
                return j
        return -1

    def _find_in_col(self, col, mark):
        for i in range(self.n):
            if self.marked[i][col] == mark:
                return i
        return -1

    def _clear_covers(self):
        for i in range(self.n):
            self.row_covered[i] = False
            self.col_covered[i] = False
