"""On-device synthetic data: the device equivalent of gen_syn_data.py:12-53.

    A ~ N(0,1)^{m x d} with unit-norm columns, Z = Bern(p)*N(mu,sigma), E likewise, X = A Z + E.

Layout follows the model, not the .mat file: Z (d,B), E (m,B), X (m,B), batch contiguous.  ``save_mat``
writes the reference's file format (keys A, train_x, test_x, train_z, test_z, train_e, test_e stored
sample-major, float64) so the reference's drivers can read it.
"""
import collections
import ctypes as C

import torch

from . import _lib

SynData = collections.namedtuple("SynData", ["A", "X", "Z", "E"])


AMPLITUDES = {"gaussian": 0, "cosine": 1, "logistic": 2}


def gen_syn_data(B, m=250, d=500, p=0.1, sigma=1.0, mu=0.0, seed=1126, device=None, A=None, col_offset=0,
                 dense_noise_sigma=None, amplitude="gaussian"):
    """Generate ``B`` problem instances on ``device``.  The value of column c depends only on
    (seed, col_offset + c), so shards generated on different GPUs tile one global data set.
    ``A`` given: reuse it (gen_syn_unseen_data.py); ``dense_noise_sigma``: E ~ N(0, s) dense
    (gen_syn_unseen_data_lasso.py:41-42); ``amplitude``: "gaussian" | "cosine" | "logistic" non-zero values
    g, cos(g), 1/(1+exp(g)) with g ~ N(mu, sigma) (gen_syn_unseen_data_cosine.py / _logistic.py:33-38)."""
    if amplitude not in AMPLITUDES:
        raise ValueError("amplitude must be one of %s" % sorted(AMPLITUDES))
    lib = _lib.load()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("gen_syn_data runs on a CUDA device only")
    g = _lib.GenDesc()
    g.m, g.d, g.B, g.col_offset, g.seed = m, d, B, col_offset, seed
    g.p, g.mu, g.sigma = p, mu, sigma
    g.dense_noise = 1 if dense_noise_sigma is not None else 0
    g.sigma_e = float(dense_noise_sigma or 0.0)
    g.amplitude = AMPLITUDES[amplitude]
    if A is None:
        A = torch.empty((m, d), dtype=torch.float32, device=device)
        g.generate_A = 1
    else:
        A = A.detach().to(device, torch.float32).contiguous()
        if tuple(A.shape) != (m, d):
            raise RuntimeError("A must be (%d,%d)" % (m, d))
        g.generate_A = 0
    Z = torch.empty((d, B), dtype=torch.float32, device=device)
    E = torch.empty((m, B), dtype=torch.float32, device=device)
    X = torch.empty((m, B), dtype=torch.float32, device=device)
    nbytes = lib.dladmm_gen_workspace_bytes(m, d)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
    g.A, g.Zs, g.Es, g.X = A.data_ptr(), Z.data_ptr(), E.data_ptr(), X.data_ptr()
    g.workspace, g.workspace_bytes = ws.data_ptr(), nbytes
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device)
        _lib.check(lib.dladmm_gen_syn(C.byref(g), stream.cuda_stream))
        ws.record_stream(stream)
    return SynData(A, X, Z, E)


def replace_A_columns(A, c, seed=19950118):
    """gen_syn_unseen_data_Acols.py:14-26: replace `c` randomly chosen columns of A by fresh unit-norm Gaussian columns
    (the "unseen dictionary" experiments).  Returns (A_new, changed_column_indices); A itself is not modified."""
    m, d = A.shape
    if not 0 <= c <= d:
        raise ValueError("c must be in [0, %d]" % d)
    g = torch.Generator().manual_seed(int(seed))
    cols = torch.randperm(d, generator=g)[:c]
    new = torch.randn(m, c, generator=g)
    new = new / new.pow(2).sum(dim=0, keepdim=True).sqrt()
    out = A.detach().clone()
    out[:, cols.to(out.device)] = new.to(out.device, out.dtype)
    return out, cols


def load_mat(path, device=None):
    """Read a reference-format syn_data .mat (gen_syn_data.py:49-53: sample-major float arrays) into the model's layout:
    returns (train, test) SynData with Z (d,B), E (m,B), X (m,B) float32, on `device` if given."""
    import scipy.io as sio
    z = sio.loadmat(path)
    A = torch.from_numpy(z["A"].astype("float32"))
    f = lambda k: torch.from_numpy(z[k].astype("float32")).t().contiguous()
    to = (lambda t: t.to(device)) if device is not None else (lambda t: t)
    sets = []
    for split in ("train", "test"):
        sets.append(SynData(to(A), to(f(split + "_x")), to(f(split + "_z")), to(f(split + "_e"))))
    return sets[0], sets[1]


def save_mat(path, train, test):
    """Write the reference's syn_data .mat layout (gen_syn_data.py:49-53) from two SynData sets that
    share A."""
    import scipy.io as sio
    t = lambda x: x.t().double().cpu().numpy()
    sio.savemat(path, dict(A=train.A.double().cpu().numpy(), train_x=t(train.X), test_x=t(test.X),
                           train_z=t(train.Z), test_z=t(test.Z), train_e=t(train.E), test_e=t(test.E)))
