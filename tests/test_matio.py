"""Matrix persistence: the reference's dense text dumps and the sparse .npz form, on the host."""
import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sparse

import psfem_oracle as O


@pytest.fixture()
def K_small():
    nodes, conn = O.mesh(4.0, 1.0, 8, 3)
    return O.global_stiffness_matrix(nodes, conn, 210e9, 0.3, 0.1)          # structural CSR, explicit zeros


@pytest.mark.parametrize("ext,sep", [(".txt", "\t"), (".csv", ",")])
def test_dense_text_round_trip_with_the_reference_recipe(psfem, tmp_path, K_small, ext, sep):
    dense = K_small.toarray()
    # written here, read the way main.py:48-49 / freeModes.py:49-50 do
    ours = tmp_path / ("ours" + ext)
    psfem.save_matrix(str(ours), K_small)
    back = pd.read_csv(ours, sep=sep, header=None, float_precision="round_trip").to_numpy()
    assert back.dtype == np.float64 and np.array_equal(back, dense)
    # (pandas' DEFAULT float parser, which the reference's scripts use, is a fast strtod that is up to 1 ulp off)
    fast = pd.read_csv(ours, sep=sep, header=None).to_numpy()
    assert np.abs(fast - dense).max() <= 4 * np.finfo(float).eps * np.abs(dense).max()
    # written the way ComputeMatrices.py:34-35 / Beamexemple.py:53-54 do, read here
    ref = tmp_path / ("ref" + ext)
    pd.DataFrame(dense).to_csv(ref, sep=sep, index=False, header=False)
    assert np.array_equal(psfem.load_matrix(str(ref)), dense)
    # the two writers agree byte for byte (repr floats, no header, no index)
    assert ours.read_bytes() == ref.read_bytes()
    # back on the structural pattern: explicit zeros (exact cancellations) are kept
    S = psfem.load_matrix(str(ref), pattern=K_small)
    assert np.array_equal(S.indptr, K_small.indptr) and np.array_equal(S.indices, K_small.indices)
    assert np.array_equal(S.data, K_small.data)
    assert psfem.load_matrix(str(ref), as_csr=True).nnz == np.count_nonzero(dense) < K_small.nnz


def test_npz_keeps_the_structural_pattern(psfem, tmp_path, K_small):
    f = tmp_path / "K.npz"
    psfem.save_matrix(str(f), K_small)
    A = psfem.load_matrix(str(f))
    assert sparse.isspmatrix_csr(A) or A.format == "csr"
    assert np.array_equal(A.indptr, K_small.indptr) and np.array_equal(A.indices, K_small.indices) and np.array_equal(A.data, K_small.data)


def test_errors(psfem, tmp_path, K_small):
    with pytest.raises(ValueError, match="cannot infer the format"):
        psfem.save_matrix(str(tmp_path / "K.bin"), K_small)
    with pytest.raises(ValueError, match="not sensible"):
        psfem.save_matrix(str(tmp_path / "K.txt"), sparse.identity(psfem.matio.DENSE_TEXT_LIMIT + 1, format="csr"))
    other = sparse.identity(K_small.shape[0], format="csr")
    f = tmp_path / "K.txt"
    psfem.save_matrix(str(f), K_small)
    with pytest.raises(ValueError, match="outside the given pattern"):
        psfem.load_matrix(str(f), pattern=other)
