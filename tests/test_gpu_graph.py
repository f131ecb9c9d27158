"""N2: sparse_batch_matmul (utils/matrix_utils.py:22-33) on the CSR SpMM kernel vs the reference expression itself
(torch.sparse.mm on the transposed / reshaped dense operand, run on CPU in fp64), values and gradients, <= 1e-5 relative."""
import numpy as np
import pytest
import torch

from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _reference_sparse_batch_matmul(sparse_matrix, dense_matrix_batch):
    b, n, p = dense_matrix_batch.shape
    dense_matrix = dense_matrix_batch.transpose(0, 1).reshape(n, b * p)
    result = torch.sparse.mm(sparse_matrix, dense_matrix)
    return result.reshape(-1, b, p).transpose(0, 1)


@pytest.mark.parametrize("res,B,p", [(8, 2, 256), (10, 3, 37), (6, 1, 4)])
def test_sparse_batch_matmul_on_the_vertex_adjacency(res, B, p):
    from deftet_b200 import builders, graph
    from deftet_b200.grid import acute_lattice_grid
    g = acute_lattice_grid(res)
    tet = torch.from_numpy(g.tets).cuda()
    adj = builders.tet_to_adj_sparse(g.n_vert, tet, normalize=True)                    # the trainer's point_adj_sparse (train_multigpu.py:77-82)
    gen = torch.Generator().manual_seed(res)
    x = torch.randn(B, g.n_vert, p, generator=gen)
    gout = torch.randn(B, g.n_vert, p, generator=gen)
    dx = x.cuda().requires_grad_(True)
    out = graph.sparse_batch_matmul(adj, dx)
    assert out.shape == (B, g.n_vert, p)
    (out * gout.cuda()).sum().backward()
    rx = x.double().requires_grad_(True)
    ref = _reference_sparse_batch_matmul(adj.cpu().double().coalesce(), rx)
    (ref * gout.double()).sum().backward()
    assert rel_err(out.detach(), ref.detach()) < 1e-5 and rel_err(dx.grad, rx.grad) < 1e-5
    assert graph.csr_of(adj) is graph.csr_of(adj)                                      # converted once per adjacency
    csr = graph.adjacency_csr(tet, g.n_vert, normalize=True)                           # straight from the tet list
    assert rel_err(graph.sparse_batch_matmul(csr, dx.detach()), ref.detach()) < 1e-5


def test_sparse_batch_matmul_rectangular_unsorted_and_empty_rows():
    from deftet_b200 import graph
    gen = torch.Generator().manual_seed(0)
    m, n, nnz, B, p = 50, 70, 400, 2, 12
    flat = torch.randperm(m * n, generator=gen)[:nnz]
    rows, cols = flat // n, flat % n
    rows[rows == 7] = 8                                                                # row 7 stays empty (duplicates are summed by coalesce)
    vals = torch.randn(nnz, generator=gen)
    sp = torch.sparse_coo_tensor(torch.stack([rows, cols]), vals, (m, n)).coalesce()
    x = torch.randn(B, n, p, generator=gen)
    dx = x.cuda().requires_grad_(True)
    out = graph.sparse_batch_matmul(sp.cuda(), dx)
    gout = torch.randn(B, m, p, generator=gen)
    (out * gout.cuda()).sum().backward()
    rx = x.double().requires_grad_(True)
    ref = _reference_sparse_batch_matmul(sp.double(), rx)
    (ref * gout.double()).sum().backward()
    assert out.shape == (B, m, p) and float(out[:, 7].abs().max()) == 0.0
    assert rel_err(out.detach(), ref.detach()) < 1e-5 and rel_err(dx.grad, rx.grad) < 1e-5
    csr = graph.csr_of(sp.cuda())
    rp = csr.fwd[0].cpu().numpy()
    assert rp[0] == 0 and rp[-1] == sp._nnz() and np.all(np.diff(rp) >= 0)
    empty = torch.sparse_coo_tensor(torch.zeros(2, 0, dtype=torch.long), torch.zeros(0), (5, 6)).cuda()
    assert float(graph.sparse_batch_matmul(empty, torch.ones(1, 6, 3).cuda()).abs().max()) == 0.0
