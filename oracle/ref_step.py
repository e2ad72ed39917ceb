"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch, CPU tensors) of the
reference's training / inference arithmetic.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this file.

Every function cites the reference lines (relative to /root/reference) it follows.
Pinned against the unmodified reference (run under oracle/dgl_shim.py) by
oracle/make_golden.py -> tests/golden/*.npz and tests/test_oracle.py.
The DGL GraphConv arithmetic itself is "parity unpinned" (third-party, not
vendored, no reference test) -- it is restated from the
published DGL 2.0.0 algorithm and self-checked against the dense form.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


# ----------------------------------------------------------------------------
# graph containers
# ----------------------------------------------------------------------------
@dataclass
class HostCSR:
    """Directed in-edge CSR of an undirected graph (both directions stored).
    dgl.from_networkx semantics: graphExtender.py:102-103 (nnz = 2|E|)."""
    rowptr: np.ndarray          # int32 [n+1]
    colidx: np.ndarray          # int32 [nnz]
    weights: np.ndarray         # float32 [nnz]  (nx edge attr 'weight', GraphCreator.py:88-90)
    n: int

    @property
    def nnz(self) -> int:
        return int(self.colidx.shape[0])

    def degrees(self) -> np.ndarray:
        return np.diff(self.rowptr).astype(np.int64)


def csr_from_networkx(nx_graph) -> HostCSR:
    """Sorted-node relabel + symmetrise, as dgl.from_networkx does for nx.Graph."""
    nodes = sorted(nx_graph.nodes())
    index = {u: i for i, u in enumerate(nodes)}
    n = len(nodes)
    rows: List[List[Tuple[int, float]]] = [[] for _ in range(n)]
    for u, v, data in nx_graph.edges(data=True):
        w = float(data.get("weight", 1))
        iu, iv = index[u], index[v]
        rows[iv].append((iu, w))
        if iu != iv:
            rows[iu].append((iv, w))
    rowptr = np.zeros(n + 1, dtype=np.int32)
    col: List[int] = []
    wts: List[float] = []
    for i, r in enumerate(rows):
        r.sort()
        rowptr[i + 1] = rowptr[i] + len(r)
        col.extend(c for c, _ in r)
        wts.extend(w for _, w in r)
    return HostCSR(rowptr, np.asarray(col, dtype=np.int32), np.asarray(wts, dtype=np.float32), n)


def dense_adjacency(csr: HostCSR, width: Optional[int] = None, dtype=torch.float32) -> torch.Tensor:
    """commons.py:38-77 (gen_adj_matrix + qubo_dict_to_torch) followed by
    graphExtender.py:28-48 (zero-pad columns to `width`)."""
    width = csr.n if width is None else width
    if width < csr.n:
        raise ValueError("N should be greater than or equal to the original matrix size.")
    A = torch.zeros((csr.n, width), dtype=dtype)
    rows = np.repeat(np.arange(csr.n), np.diff(csr.rowptr))
    A[torch.from_numpy(rows), torch.from_numpy(csr.colidx.astype(np.int64))] = \
        torch.from_numpy(csr.weights).to(dtype)
    return A


# ----------------------------------------------------------------------------
# model arithmetic
# ----------------------------------------------------------------------------
def _aggregate(csr: HostCSR, h: torch.Tensor, use_weights: bool = False) -> torch.Tensor:
    """DGL update_all(copy_u, sum): y_v = sum_{u->v} h_u (edge weights are NOT
    used by the reference's GraphConv call, TrainingNeural.py:80,83)."""
    rows = torch.from_numpy(np.repeat(np.arange(csr.n), np.diff(csr.rowptr)))
    cols = torch.from_numpy(csr.colidx.astype(np.int64))
    msg = h.index_select(0, cols)
    if use_weights:
        msg = msg * torch.from_numpy(csr.weights).to(h.dtype).unsqueeze(1)
    out = torch.zeros_like(h)
    out.index_add_(0, rows, msg)
    return out


def graphconv(csr: HostCSR, X: torch.Tensor, W: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    """dgl.nn.pytorch.GraphConv.forward, norm='both' (call sites
    TrainingNeural.py:80 and :83; semantics in oracle/dgl_shim.py docstring)."""
    deg = torch.from_numpy(csr.degrees())
    if bool((deg == 0).any()):
        raise ValueError("There are 0-in-degree nodes in the graph")
    norm = deg.to(X.dtype).clamp(min=1).pow(-0.5).unsqueeze(1)
    h = X * norm
    if W.shape[0] > W.shape[1]:
        h = _aggregate(csr, h @ W)
    else:
        h = _aggregate(csr, h) @ W
    h = h * norm
    if b is not None:
        h = h + b
    return h


@dataclass
class GCNParams:
    """GCNSoftmax parameters, TrainingNeural.py:72-77 (state_dict keys conv{1,2}.{weight,bias})."""
    W1: torch.Tensor
    b1: torch.Tensor
    W2: torch.Tensor
    b2: torch.Tensor

    def tensors(self) -> List[torch.Tensor]:
        return [self.W1, self.b1, self.W2, self.b2]

    def clone(self) -> "GCNParams":
        return GCNParams(*[t.detach().clone() for t in self.tensors()])

    @staticmethod
    def init(F: int, H: int, K: int, seed: int = 0, dtype=torch.float32) -> "GCNParams":
        """xavier_uniform_ weights, zero bias (DGL GraphConv.reset_parameters)."""
        g = torch.Generator().manual_seed(seed)
        def xavier(i, o):
            a = math.sqrt(6.0 / (i + o))
            return ((torch.rand((i, o), generator=g, dtype=torch.float64) * 2 - 1) * a).to(dtype)
        return GCNParams(xavier(F, H), torch.zeros(H, dtype=dtype), xavier(H, K), torch.zeros(K, dtype=dtype))


def gcn_forward(csr: HostCSR, X: torch.Tensor, p: GCNParams) -> Dict[str, torch.Tensor]:
    """GCNSoftmax.forward, TrainingNeural.py:79-85 (dropout p=0.0 is the identity, :43)."""
    pre1 = graphconv(csr, X, p.W1, p.b1)
    H1 = torch.relu(pre1)
    Z = graphconv(csr, H1, p.W2, p.b2)
    P = torch.softmax(Z, dim=1)
    return {"pre1": pre1, "H1": H1, "Z": Z, "P": P}


def _bf16(t: torch.Tensor) -> torch.Tensor:
    """Round to nearest even bf16 and back (what cvt.rn.bf16x2.f32 / torch's .to(bfloat16) do), keeping the dtype."""
    return t.to(torch.bfloat16).to(t.dtype)


def gcn_forward_bf16_storage(csr: HostCSR, X: torch.Tensor, p: GCNParams, preaggregated: bool = True) -> Dict[str, torch.Tensor]:
    """GCNSoftmax.forward (TrainingNeural.py:79-85) with the STORAGE roundings of the B200 throughput configuration
    applied at exactly the points where its kernels store bf16, all arithmetic in the dtype of X (use float64):
      preaggregated: XA = bf16(A_hat X), W1b = bf16(W1), H1 = bf16(relu(XA W1b + b1)), T2 = H1 W2, Z = A_hat T2 + b2
      standard     : Xb = bf16(X), W1b = bf16(W1), T1 = bf16(Xb W1b), H1 = bf16(relu(A_hat T1 + b1)), T2 = H1 W2, ...
    TEST INFRASTRUCTURE: lets the GPU tests pin the bf16 paths to ~1e-3 instead of comparing them with fp32 at 1e-2."""
    dt = X.dtype
    deg = torch.from_numpy(csr.degrees())
    norm = deg.to(dt).clamp(min=1).pow(-0.5).unsqueeze(1)

    def ahat(h):
        return _aggregate(csr, h * norm) * norm

    W1b = _bf16(p.W1.to(dt))
    if preaggregated:
        XA = _bf16(ahat(X))
        pre1 = XA @ W1b + p.b1.to(dt)
    else:
        T1 = _bf16(_bf16(X) @ W1b)
        pre1 = ahat(T1) + p.b1.to(dt)
    H1 = _bf16(torch.relu(pre1))
    T2 = H1 @ p.W2.to(dt)
    Z = ahat(T2) + p.b2.to(dt)
    return {"pre1": pre1, "H1": H1, "Z": Z, "P": torch.softmax(Z, dim=1)}


def hard_labels(P: torch.Tensor, override_terminals: bool = True) -> torch.Tensor:
    """argmax rows (first max on ties, torch.argmax; TrainingNeural.py:96-106) with
    rows 0,1,2 forced to classes 0,1,2 (override_fixed_nodes, :87-94)."""
    lab = torch.argmax(P, dim=1)
    if override_terminals:
        k = min(3, P.shape[0])
        lab[:k] = torch.arange(k)
    return lab


def cut_of_labels(csr: HostCSR, labels) -> float:
    """calculate_HC_vectorized on one-hot s (TrainingNeural.py:154-176):
    sum_ij A_ij (1 - [l_i == l_j]) / 2."""
    lab = np.asarray(labels)
    rows = np.repeat(np.arange(csr.n), np.diff(csr.rowptr))
    diff = lab[rows] != lab[csr.colidx]
    return float(np.sum(csr.weights[diff].astype(np.float64)) / 2.0)


def ste_loss_and_grads(csr: HostCSR, P: torch.Tensor, C: float = 1.0, override_terminals: bool = True,
                       mode: str = "ste", penalty: float = 0.0) -> Dict[str, torch.Tensor]:
    """Closed form of what autograd yields for
        override_fixed_nodes -> apply_max_to_one_hot -> compute_loss   (TrainingNeural.py:373-380)
    loss = -C * cut(hard labels);  g = dL/ds = C * (A s) with s the hard one-hot rows
    (STE :101 and override :91-93 are identity for gradients, all rows incl. 0-2);
    dZ = P * (g - <P, g>) with the raw softmax output P.

    mode='soft' is the north-star objective: s = P (terminal rows replaced by
    one-hots when override_terminals), loss = -C/2 sum_ij A_ij (1 - s_i.s_j),
    plus penalty * sum_{i<j in T} s_i.s_j (TrainingNeural.py:178-195, disabled at :308)."""
    n, K = P.shape
    if mode == "ste":
        lab = hard_labels(P, override_terminals)
        S = torch.zeros_like(P)
        S[torch.arange(n), lab] = 1.0
    else:
        S = P.clone()
        if override_terminals:
            k = min(3, n)
            S[:k] = torch.eye(K, dtype=P.dtype)[:k]
    AS = _aggregate(csr, S, use_weights=True)
    wdeg = _aggregate(csr, torch.ones((n, 1), dtype=P.dtype), use_weights=True).squeeze(1)
    cut = 0.5 * (wdeg.sum() - (S * AS).sum())
    loss = -C * cut
    g = C * AS
    if penalty != 0.0:
        k = min(3, n)
        T = S[:k]
        tot = T.sum(0, keepdim=True)
        loss = loss + penalty * 0.5 * ((tot * tot).sum() - (T * T).sum())
        g = g.clone()
        g[:k] += penalty * (tot - T)
    dZ = P * (g - (P * g).sum(1, keepdim=True))
    return {"loss": loss, "g": g, "dZ": dZ, "S": S}


def gcn_backward(csr: HostCSR, X: torch.Tensor, p: GCNParams, fwd: Dict[str, torch.Tensor],
                 dZ: torch.Tensor, need_dX: bool = False) -> Dict[str, torch.Tensor]:
    """Manual reverse pass of GCNSoftmax (what loss.backward(), TrainingNeural.py:385, computes).
    A_hat = D^-1/2 A D^-1/2 is symmetric, so the transposed SpMM is the same SpMM."""
    deg = torch.from_numpy(csr.degrees())
    norm = deg.to(X.dtype).clamp(min=1).pow(-0.5).unsqueeze(1)

    def ahat(h):
        return _aggregate(csr, h * norm) * norm

    db2 = dZ.sum(0)
    dT2 = ahat(dZ)                      # grad wrt (H1 W2)
    dW2 = fwd["H1"].t() @ dT2
    dH1 = dT2 @ p.W2.t()
    dpre1 = dH1 * (fwd["pre1"] > 0).to(dH1.dtype)
    db1 = dpre1.sum(0)
    dT1 = ahat(dpre1)                   # grad wrt (X W1)
    dW1 = X.t() @ dT1
    out = {"W1": dW1, "b1": db1, "W2": dW2, "b2": db2}
    if need_dX:
        out["X"] = dT1 @ p.W1.t()
    return out


# ----------------------------------------------------------------------------
# Adam (torch.optim.Adam defaults; TrainingNeural.py:336-337)
# ----------------------------------------------------------------------------
@dataclass
class AdamState:
    step: int = 0
    m: List[torch.Tensor] = field(default_factory=list)
    v: List[torch.Tensor] = field(default_factory=list)


def adam_update(params: Sequence[torch.Tensor], grads: Sequence[torch.Tensor], st: AdamState,
                lr: float = 1e-3, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8) -> None:
    """torch.optim.Adam single-tensor formula (amsgrad=False, weight_decay=0, maximize=False)."""
    if not st.m:
        st.m = [torch.zeros_like(p) for p in params]
        st.v = [torch.zeros_like(p) for p in params]
    st.step += 1
    bc1 = 1.0 - beta1 ** st.step
    bc2 = 1.0 - beta2 ** st.step
    step_size = lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    for p, g, m, v in zip(params, grads, st.m, st.v):
        m.lerp_(g, 1.0 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
        denom = (v.sqrt() / bc2_sqrt).add_(eps)
        p.addcdiv_(m, denom, value=-step_size)


# ----------------------------------------------------------------------------
# whole steps
# ----------------------------------------------------------------------------
def train_step_closed_form(csr: HostCSR, X: torch.Tensor, p: GCNParams, st: AdamState,
                           lr: float = 1e-3, C: float = 1.0, mode: str = "ste",
                           penalty: float = 0.0) -> float:
    """One reference optimiser step for one graph (TrainingNeural.py:371-388), vectorised."""
    fwd = gcn_forward(csr, X, p)
    lg = ste_loss_and_grads(csr, fwd["P"], C, True, mode, penalty)
    grads = gcn_backward(csr, X, p, fwd, lg["dZ"])
    adam_update(p.tensors(), [grads["W1"], grads["b1"], grads["W2"], grads["b2"]], st, lr)
    return float(lg["loss"])


def batch_loss_and_grads(graphs: Sequence[Tuple[HostCSR, torch.Tensor]], p: GCNParams, C: float = 1.0,
                         mode: str = "ste", penalty: float = 0.0):
    """Sum over graphs of per-graph loss / weight grads at FIXED weights -- the
    batched / data-parallel parity definition of SURVEY.md 8(e)."""
    tot = None
    losses = []
    for csr, X in graphs:
        fwd = gcn_forward(csr, X, p)
        lg = ste_loss_and_grads(csr, fwd["P"], C, True, mode, penalty)
        g = gcn_backward(csr, X, p, fwd, lg["dZ"])
        losses.append(float(lg["loss"]))
        tot = g if tot is None else {k: tot[k] + g[k] for k in g}
    return losses, tot


class FaithfulPort:
    """The reference's per-graph step with the SAME cost structure: dense
    [n, n_pad] features through two matmuls, a Python loop over rows for the
    straight-through one-hot, a dense [n, n_pad] loss and torch autograd +
    torch.optim.Adam.  Used as the CPU baseline ("kind": "port") because the
    Python reference itself cannot travel to the GPU box.

    Follows TrainingNeural.py:79-106 (model, override, STE), :154-176 / :291-309
    (loss, pad width hard-coded 1000 at :171) and :371-388 (step order)."""

    def __init__(self, F: int, H: int, K: int = 3, lr: float = 1e-3, seed: int = 0, pad: int = 1000):
        init = GCNParams.init(F, H, K, seed)
        self.params = [torch.nn.Parameter(t) for t in init.tensors()]
        self.opt = torch.optim.Adam(self.params, lr=lr)
        self.pad = pad

    def _conv(self, csr, rows, cols, norm, h, W, b):
        h = (h * norm) @ W
        out = torch.zeros((csr.n, W.shape[1]), dtype=h.dtype)
        out = out.index_add(0, rows, h.index_select(0, cols))
        return out * norm + b

    def forward(self, csr: HostCSR, X: torch.Tensor) -> torch.Tensor:
        W1, b1, W2, b2 = self.params
        rows = torch.from_numpy(np.repeat(np.arange(csr.n), np.diff(csr.rowptr)))
        cols = torch.from_numpy(csr.colidx.astype(np.int64))
        norm = torch.from_numpy(csr.degrees()).to(X.dtype).clamp(min=1).pow(-0.5).unsqueeze(1)
        h = torch.relu(self._conv(csr, rows, cols, norm, X, W1, b1))
        return torch.softmax(self._conv(csr, rows, cols, norm, h, W2, b2), dim=1)

    @staticmethod
    def _hard(P: torch.Tensor) -> torch.Tensor:
        fixed = P.clone()
        eye = torch.eye(3, dtype=P.dtype)
        for t in range(3):
            fixed[t] = eye[t] + P[t] - P[t].detach()
        rows = []
        for i in range(fixed.size(0)):          # the reference's per-row Python loop (:106)
            r = fixed[i]
            oh = torch.zeros_like(r)
            oh[torch.argmax(r)] = 1.0
            rows.append(oh + r - r.detach())
        return torch.stack(rows)

    def loss(self, P: torch.Tensor, A_pad: torch.Tensor, C: float = 1.0) -> torch.Tensor:
        s = self._hard(P)
        same = s @ s.t()
        same_pad = torch.zeros(same.shape[0], self.pad)
        same_pad[:, : same.shape[0]] = same
        return C * (-(A_pad * (1 - same_pad)).sum() / 2)

    def step(self, csr: HostCSR, X: torch.Tensor, A_pad: torch.Tensor, C: float = 1.0) -> float:
        loss = self.loss(self.forward(csr, X), A_pad, C)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return loss.item()

    @torch.no_grad()
    def evaluate(self, csr: HostCSR, X: torch.Tensor, A_pad: torch.Tensor, C: float = 1.0) -> float:
        """evaluate_model body, TrainingNeural.py:553-562."""
        return self.loss(self.forward(csr, X), A_pad, C).item()
