"""TEST INFRASTRUCTURE (oracle): square linear sum assignment, restated.

The reference calls `scipy.optimize.linear_sum_assignment` on every p x p matching cost
matrix (multi_part_assembly/models/modules/base_model.py:175-176).  SciPy is a third-party
dependency that the reference pins no version of; its `rectangular_lsap` implements the
shortest-augmenting-path algorithm of D. F. Crouse, "On implementing 2D rectangular
assignment algorithms", IEEE TAES 52(4), 2016.  This is a plain-Python restatement of that
algorithm (float64 duals, columns visited in descending index order, ties resolved towards
an unassigned column), pinned against the installed SciPy by tests/test_oracle_cpu.py; the
CUDA kernel `lsap_kernel` (csrc/loss.cu) follows the same steps.
"""
import math


def linear_sum_assignment_square(cost):
    """cost: n x n nested sequence / array.  Returns col4row (list of n ints)."""
    n = len(cost)
    u = [0.0] * n
    v = [0.0] * n
    path = [-1] * n
    col4row = [-1] * n
    row4col = [-1] * n
    for cur in range(n):
        min_val = 0.0
        remaining = [n - it - 1 for it in range(n)]
        num_remaining = n
        SR = [False] * n
        SC = [False] * n
        spc = [math.inf] * n
        sink, i = -1, cur
        while sink == -1:
            index, lowest = -1, math.inf
            SR[i] = True
            for it in range(num_remaining):
                j = remaining[it]
                r = min_val + float(cost[i][j]) - u[i] - v[j]
                if r < spc[j]:
                    path[j] = i
                    spc[j] = r
                if spc[j] < lowest or (spc[j] == lowest and row4col[j] == -1):
                    lowest = spc[j]
                    index = it
            min_val = lowest
            if index < 0 or min_val == math.inf:
                return list(range(n))  # infeasible: identity (SciPy raises)
            j = remaining[index]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            SC[j] = True
            num_remaining -= 1
            remaining[index] = remaining[num_remaining]
        u[cur] += min_val
        for r in range(n):
            if SR[r] and r != cur:
                u[r] += min_val - spc[col4row[r]]
        for j in range(n):
            if SC[j]:
                v[j] -= min_val - spc[j]
        j = sink
        while True:
            r = path[j]
            row4col[j] = r
            col4row[r], j = j, col4row[r]
            if r == cur:
                break
    return col4row
