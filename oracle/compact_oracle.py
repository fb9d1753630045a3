"""CPU restatement (float64 NumPy) of the compact-state algebra of the learner's first layer — TEST INFRASTRUCTURE.

Only tests/ may import this.  It pins the exact identities the CUDA compact path relies on (csrc/dcc_compact.cuh):

The reference's observation row of agent i (envs/mpe/multiagent/scenarios/coverage.py:99-110) is an AFFINE function of
the env's compact state s = (p_k, v_k for the N UAVs; energy_j for the M PoIs):

    x = [ v_i (2), p_i (2), p_k - p_i for k != i (2(N-1)),  then per PoI j:  q_j - p_i (2), e_j, 5.0, [e_j >= 5] ]

and the centralised critic input is the N rows of the env concatenated (learner.py:219-220).  With the input LayerNorm
(algos/algo_utils/mlp.py:44-58) written as xhat = rstd * x - rstd * mean, every xhat is LINEAR in the feature vector

    f = rstd * [ own_0 .. own_{nb-1} (nb * OWN),  e_1..e_M,  d_1..d_M,  1,  -mean ]        (OWN = 2N + 2)

(nb = 1 block for an actor row, N blocks for a critic row; own_i = the first OWN entries of agent i's row), i.e.
xhat = A f with a fixed sparse matrix A that depends only on the PoI table.  Hence, exactly:

    forward        xhat W^T = f (W A)^T                    -> `fold_weights`  (K shrinks 338 -> 148 / 2704 -> 274 at 8/64)
    weight grad    dz^T xhat = (dz^T f) A^T                -> `unfold_grad`
"""
import numpy as np

LN_EPS = 1e-5


def obs_rows(pos_vel, energy, poi, m_energy=5.0):
    """Observation rows exactly as the reference builds them (coverage.py:99-110), float32 of float64 expressions.
    pos_vel (R, N, 4) float64 [px, py, vx, vy]; energy (R, M); poi (M, 2) -> (R, N, D) float32."""
    pv = np.asarray(pos_vel, dtype=np.float64)
    R, N, _ = pv.shape
    en = np.asarray(energy, dtype=np.float64).reshape(R, -1)
    M = en.shape[1]
    D = 4 + 2 * (N - 1) + 5 * M
    out = np.zeros((R, N, D), dtype=np.float32)
    p, v = pv[:, :, 0:2], pv[:, :, 2:4]
    for i in range(N):
        out[:, i, 0:2] = v[:, i]
        out[:, i, 2:4] = p[:, i]
        others = [k for k in range(N) if k != i]
        if others:
            out[:, i, 4:4 + 2 * (N - 1)] = (p[:, others] - p[:, i:i + 1]).reshape(R, -1)
        base = 2 * N + 2
        blk = np.zeros((R, M, 5))
        blk[:, :, 0:2] = poi[None] - p[:, i:i + 1]
        blk[:, :, 2] = en
        blk[:, :, 3] = m_energy
        blk[:, :, 4] = (en >= m_energy)
        out[:, i, base:] = blk.reshape(R, -1)
    return out


def feature_dims(N, M, nb):
    own = 2 * N + 2
    return own, nb * own + 2 * M + 2


def features(pos_vel, energy, poi, centralized, normalize=True, m_energy=5.0):
    """Feature rows f (float64).  centralized=False: one row per agent (R*N, K_a); True: one row per env (R, K_c).
    The LayerNorm statistics are those of the float32 observation row the reference would normalise."""
    x = obs_rows(pos_vel, energy, poi, m_energy).astype(np.float64)
    R, N, D = x.shape
    M = np.asarray(energy).reshape(R, -1).shape[1]
    own = 2 * N + 2
    en = x[:, 0, own + 2::5]
    dn = x[:, 0, own + 4::5]
    if centralized:
        rows = x.reshape(R, N * D)
        blocks = x[:, :, :own].reshape(R, N * own)
    else:
        rows = x.reshape(R * N, D)
        blocks = x[:, :, :own].reshape(R * N, own)
        en, dn = np.repeat(en, N, axis=0), np.repeat(dn, N, axis=0)
    if normalize:
        mean = rows.mean(axis=1, keepdims=True)
        rstd = 1.0 / np.sqrt(rows.var(axis=1, keepdims=True) + LN_EPS)
    else:
        mean, rstd = np.zeros((rows.shape[0], 1)), np.ones((rows.shape[0], 1))
    f = np.concatenate([blocks, en, dn, np.ones_like(mean), -mean], axis=1) * rstd
    return f


def fold_weights(Wg, poi, N, M, nb, m_energy=5.0):
    """Wg (H, nb*D) [fc1 weight with the LayerNorm gain folded in] -> Wt (H, K) with xhat Wg^T == f Wt^T."""
    own, K = feature_dims(N, M, nb)
    D = 4 + 2 * (N - 1) + 5 * M
    H = Wg.shape[0]
    Wt = np.zeros((H, K))
    for i in range(nb):
        blk = Wg[:, i * D:(i + 1) * D]
        Wt[:, i * own:(i + 1) * own] = blk[:, :own]
        Wt[:, i * own + 2] -= blk[:, own + 0::5].sum(axis=1)
        Wt[:, i * own + 3] -= blk[:, own + 1::5].sum(axis=1)
        Wt[:, nb * own:nb * own + M] += blk[:, own + 2::5]
        Wt[:, nb * own + M:nb * own + 2 * M] += blk[:, own + 4::5]
        Wt[:, nb * own + 2 * M] += blk[:, own + 0::5] @ poi[:, 0] + blk[:, own + 1::5] @ poi[:, 1] + \
            m_energy * blk[:, own + 3::5].sum(axis=1)
    Wt[:, nb * own + 2 * M + 1] = Wg.sum(axis=1)
    return Wt


def unfold_grad(Gt, poi, N, M, nb, m_energy=5.0):
    """Gt (H, K) = dz^T f  ->  G (H, nb*D) = dz^T xhat."""
    own, K = feature_dims(N, M, nb)
    D = 4 + 2 * (N - 1) + 5 * M
    H = Gt.shape[0]
    G = np.zeros((H, nb * D))
    gc, gm = Gt[:, nb * own + 2 * M], Gt[:, nb * own + 2 * M + 1]
    for i in range(nb):
        blk = G[:, i * D:(i + 1) * D]
        blk[:, :own] = Gt[:, i * own:(i + 1) * own]
        blk[:, own + 0::5] = gc[:, None] * poi[None, :, 0] - Gt[:, i * own + 2][:, None]
        blk[:, own + 1::5] = gc[:, None] * poi[None, :, 1] - Gt[:, i * own + 3][:, None]
        blk[:, own + 2::5] = Gt[:, nb * own:nb * own + M]
        blk[:, own + 3::5] = m_energy * gc[:, None]
        blk[:, own + 4::5] = Gt[:, nb * own + M:nb * own + 2 * M]
    return G + gm[:, None]


def state_from_obs(obs, N, M):
    """Inverse of obs_rows on recorded float32 observations (golden rollouts): positions / velocities as float64 of the
    stored float32 values, energies as integers.  obs (..., N, D) -> pos_vel (..., N, 4), energy (..., M) uint8."""
    o = np.asarray(obs)
    own = 2 * N + 2
    pv = np.concatenate([o[..., 2:4], o[..., 0:2]], axis=-1).astype(np.float64)
    en = np.rint(o[..., 0, own + 2::5]).astype(np.uint8)
    return pv, en
