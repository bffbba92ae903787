"""Replay backoff policies of the warm-started split on the table written by tools/refusal_study.py (CPU only).

For every (bond, direction) the table says, visit by visit, what the device-side gates WOULD answer if the fast path were
attempted.  A policy decides at each visit whether to attempt; an attempt that is refused costs `c_ref` ms more than an
accepted one (fast attempt + single-CTA cold pipeline on the critical path), a visit that sits out costs `c_cold` ms more
(cold cluster pipeline instead of the fast split).  Cost models: one GPU c_ref = 2.1, c_cold = 0.05
(the projection hides most of either split); eight GPUs c_ref = 2.45, c_cold = 0.33.

    python tools/refusal_policy_eval.py /tmp/refusal_table.npy [first_timed_sweep]
"""
import sys

import numpy as np

COSTS = {"1 GPU": (2.1, 0.05), "8 GPUs": (2.45, 0.33)}


def backoff(free_retries, waits):
    """-> policy(state, feat_prev) ; state = dict(fails, wait).  waits[k] = visits to sit out after the k-th refusal in a
    row beyond the free ones (last entry repeats)."""
    def decide(st, feat_prev):
        if st.get("wait", 0) > 0:
            st["wait"] -= 1
            return False
        return True

    def feedback(st, accepted):
        if accepted:
            st["fails"] = 0
            return
        n = st["fails"] = st.get("fails", 0) + 1
        if n > free_retries:
            st["wait"] = waits[min(n - free_retries - 1, len(waits) - 1)]
    return decide, feedback


def spectrum(max_tau, max_ratio, inner):
    """Attempt only when the PREVIOUS visit's spectrum (known to the host from the sweep's record, whichever pipeline
    ran) shows the gap the gates ask for, with a margin; otherwise like `inner`."""
    d0, f0 = inner

    def decide(st, feat_prev):
        if feat_prev is not None and not (feat_prev[0] <= max_tau and feat_prev[1] <= max_ratio):
            if st.get("wait", 0) > 0:
                st["wait"] -= 1
            return False
        return d0(st, feat_prev)
    return decide, f0


def graded(min_ratio, inner):
    """engine.SweepEngine's opt-in rule (TNML_FAST_MIN_RATIO): no attempt while the kept singular values of the bond were
    graded below (sigma_m / sigma_1)^2 = min_ratio at its previous visit; otherwise like `inner`."""
    d0, f0 = inner

    def decide(st, feat_prev):
        go = d0(st, feat_prev)                            # (the wait counter runs on)
        return go and not (feat_prev is not None and feat_prev[2] < min_ratio)
    return decide, f0


def clairvoyant():
    return "clairvoyant", None


POLICIES = {
    "always attempt": backoff(10 ** 9, [0]),
    "round 2 before (1 free retry, then 2)": backoff(1, [2]),
    "exponential (1 free retry, then 4, 8, 16)": backoff(1, [4, 8, 16]),
    "no free retry: 4, 8, 16": backoff(0, [4, 8, 16]),
    "no free retry: 2, 4, 8, 16": backoff(0, [2, 4, 8, 16]),
    "no free retry: 8, 16, 32": backoff(0, [8, 16, 32]),
    "spectrum tau<0.1 + exponential": spectrum(0.1, 1.0, backoff(1, [4, 8, 16])),
    "spectrum tau<0.05 + exponential": spectrum(0.05, 1.0, backoff(1, [4, 8, 16])),
    "spectrum tau<0.02 + no free retry 4,8,16": spectrum(0.02, 1.0, backoff(0, [4, 8, 16])),
    "kept ratio >= 0.1 + exponential": graded(0.1, backoff(1, [4, 8, 16])),
    "kept ratio >= 0.01 + exponential": graded(0.01, backoff(1, [4, 8, 16])),
    "clairvoyant (attempt iff accepted)": clairvoyant(),
    "never attempt": (lambda st, fp: False, lambda st, a: None),
}


def replay(tab, policy, first_timed):
    keys = {}
    for r in tab:
        keys.setdefault((int(r[1]), int(r[2])), []).append(r)
    sweeps = sorted(set(int(s) for s in tab[:, 0]))
    per = {s: [0, 0, 0] for s in sweeps}                 # accepted, refused, sat out
    for key, rows in keys.items():
        rows.sort(key=lambda r: r[3])
        st, prev = {}, None
        for r in rows:
            ok = r[4] == 0.0
            s = int(r[0])
            if policy[0] == "clairvoyant":
                att = ok
            else:
                att = policy[0](st, prev)
            if att:
                per[s][0 if ok else 1] += 1
                if policy[0] != "clairvoyant":
                    policy[1](st, ok)
            else:
                per[s][2] += 1
            prev = (r[6], r[7], r[8])                     # tau / lam_m, lam_{m+1} / lam_m, lam_m / lam_1 of THIS visit
    timed = [s for s in sweeps if s >= first_timed]
    acc = sum(per[s][0] for s in timed)
    ref = sum(per[s][1] for s in timed)
    cold = sum(per[s][2] for s in timed)
    return acc, ref, cold, len(timed), per


def main():
    tab = np.load(sys.argv[1])
    first_timed = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    sweeps = sorted(set(int(s) for s in tab[:, 0]))
    print("table: %d rows, sweeps %d..%d; timed region = sweeps >= %d" % (len(tab), sweeps[0], sweeps[-1], first_timed))
    print("\nwould-be outcome of an attempt at every revisit, per sweep (code: count):")
    for s in sweeps:
        m = tab[tab[:, 0] == s]
        print("  sweep %2d: " % s + "  ".join("%g: %d" % (c, (m[:, 4] == c).sum()) for c in np.unique(m[:, 4])))
    # how predictable is a refusal?
    print("\nP(accepted at visit v | outcome at visit v-1), visits >= 3:")
    keys = {}
    for r in tab:
        keys.setdefault((int(r[1]), int(r[2])), []).append(r)
    cnt = np.zeros((2, 2))
    for rows in keys.values():
        rows.sort(key=lambda r: r[3])
        for a, b in zip(rows[:-1], rows[1:]):
            if b[0] >= first_timed:
                cnt[int(a[4] == 0), int(b[4] == 0)] += 1
    for prev_ok in (1, 0):
        tot = cnt[prev_ok].sum()
        print("  previous %s: %d visits, accepted %.3f" % ("accepted" if prev_ok else "refused ", tot,
                                                            cnt[prev_ok, 1] / max(1, tot)))
    print("\n%-46s %8s %8s %8s   %s" % ("policy (per sweep, timed region)", "accepted", "refused", "sat out",
                                      "  ".join("extra ms/sweep " + k for k in COSTS)))
    for name, pol in POLICIES.items():
        acc, ref, cold, n, _ = replay(tab, pol, first_timed)
        costs = "  ".join("%20.2f" % ((ref * cr + cold * cc) / n) for cr, cc in COSTS.values())
        print("%-46s %8.1f %8.1f %8.1f   %s" % (name, acc / n, ref / n, cold / n, costs))


if __name__ == "__main__":
    main()
