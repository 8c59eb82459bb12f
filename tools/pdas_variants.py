"""Numpy laboratory for the working-set rule of the formulation-A kernel (csrc/forma.cuh / forma_reg.cuh): replays variants
of the structured primal-dual active-set iteration on the kernel's own dumped workload and prints their iteration
histograms.  Needs no GPU -- only the dumps `tools/forma_dump.py 1024 trot` / `2048 walk` leave in gpurun_out/.

usage: python tools/pdas_variants.py trot|walk [n_instances] [variants, e.g. 0,2,6] [theta]
  variant 0  the rule of round 1 (peeling cut only from a feasible point)
  variant 2  the rule the kernel runs now (the cut does not wait for feasibility): slowest QP 14 -> 11 iterations
  variant 1 / 3  two-phase rules (add while violated, release only when feasible): worse, some cycle
  variant 4  peel whenever a run end is wrong, interior wrong rows or not: a few instances cycle
  variant 5 / 6 / 7  violated rows enter only above 5 / 20 / 50 % of the largest violation: 6 trims the tail a little (max 10)
  theta < 1  under-relaxed cut (release only a fraction of the rows the closed form proposes): fewer 11s, same maximum
Every variant is checked against variant 0's minimiser (1e-7).  The dense solves come from tests/test_pdas_restatement.py."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import test_pdas_restatement as T
from oracle import oracle as O
DAMP_AFTER = 12

THETA = 1.0
def peel_cut(st, new, lam, x, mu0, Hd, aeq, Ain, lo, hi, C_, ends):
    Qz, a, dt = Hd[0], aeq[:C_], Ain[0, 0]
    PA = np.cumsum(a)
    nu = -mu0
    tgt = np.where(st[:C_] < 0, lo[:C_], hi[:C_]) - (Ain[:C_, C_:] @ x[C_:])
    cseg = -dt * np.cumsum(lam[:C_][::-1])[::-1]
    acts = np.nonzero(st[:C_])[0]
    good = lambda sig, y: y > 0 if sig < 0 else y < 0
    for i, (is_left, is_right) in ends.items():
        sig = st[i]
        if is_right:
            s = i
            while s - 1 >= 0 and st[s - 1] == sig: s -= 1
            nx = acts[acts > i]; kn = nx[0] if len(nx) else C_
            best = None
            for e in range(i, s - 1, -1):
                c_new = 0.0 if kn == C_ else ((Qz / dt) * (tgt[kn] - tgt[e]) - nu * (PA[kn] - PA[e])) / (kn - e)
                c_prev = (Qz / dt) * (tgt[e] - tgt[e - 1]) - nu * a[e] if e > s else cseg[s]
                if good(sig, c_prev - c_new): best = e; break
            fr = (best + 1 if best is not None else s)
            fr = i + 1 - max(1, int(np.ceil(THETA * (i + 1 - fr))))
            new[fr:i + 1] = 0
        if is_left:
            e = i
            while e + 1 < C_ and st[e + 1] == sig: e += 1
            pv = acts[acts < i]; kp = pv[-1] if len(pv) else -1
            tkp, PAkp = (tgt[kp], PA[kp]) if kp >= 0 else (0.0, 0.0)
            best = None
            for s2 in range(i, e + 1):
                c_new = ((Qz / dt) * (tgt[s2] - tkp) - nu * (PA[s2] - PAkp)) / (s2 - kp)
                c_next = (Qz / dt) * (tgt[s2 + 1] - tgt[s2]) - nu * a[s2 + 1] if s2 < e else (cseg[e + 1] if e + 1 < C_ else 0.0)
                if good(sig, c_new - c_next): best = s2; break
            to = (best if best is not None else e + 1)
            to = i + max(1, int(np.ceil(THETA * (to - i))))
            new[i:to] = 0

def pdas(Hd, g, aeq, beq, Ain, lo, hi, C_, variant=0, maxit=100, log=None):
    m = len(lo)
    st = np.zeros(m, int)
    for it in range(maxit):
        x, lam, mu0 = T._solve_on(Hd, g, aeq, beq, Ain, lo, hi, st)
        r = Ain @ x
        vlo = (lo - r) > 1e-10 * (1 + np.abs(lo)); vhi = (r - hi) > 1e-10 * (1 + np.abs(hi))
        viol = (st == 0) & (vlo | vhi)
        new = np.where(st == 0, np.where(vlo, -1, np.where(vhi, 1, 0)), st)
        wrong = (st != 0) & ~np.where(st < 0, lam < 0, lam > 0)
        nviol = int(viol.sum())
        wz = np.nonzero(wrong[:C_])[0]
        ends = {}
        for i in wz:
            sl = st[i - 1] if i > 0 else 0; sr = st[i + 1] if i + 1 < C_ else 0
            if sl != st[i] or sr != st[i]: ends[i] = (sl != st[i], sr != st[i])
        if log is not None:
            log.append(''.join('-' if v<0 else ('+' if v>0 else '.') for v in st[:C_]) + ' v%d w%d e%d'%(nviol, wrong.sum(), len(ends)))
        if variant == 0:
            new[wrong] = 0
            if ends and nviol == 0 and len(ends) == len(wz):
                peel_cut(st, new, lam, x, mu0, Hd, aeq, Ain, lo, hi, C_, ends)
            elif it >= DAMP_AFTER and ends:
                for i in wz:
                    if i not in ends: new[i] = st[i]
        elif variant == 1:
            # two-phase: while rows are violated only add; release (with peeling where it applies) only from a feasible point
            if nviol == 0:
                new[wrong] = 0
                if ends and len(ends) == len(wz):
                    peel_cut(st, new, lam, x, mu0, Hd, aeq, Ain, lo, hi, C_, ends)
        elif variant == 2:
            # peel even with violations present
            new[wrong] = 0
            if ends and len(ends) == len(wz):
                peel_cut(st, new, lam, x, mu0, Hd, aeq, Ain, lo, hi, C_, ends)
            elif it >= DAMP_AFTER and ends:
                for i in wz:
                    if i not in ends: new[i] = st[i]
        elif variant in (5, 6, 7):
            # violated rows enter only if their violation is at least FRAC of the largest one
            FR = {5: 0.05, 6: 0.2, 7: 0.5}[variant]
            vv = np.maximum(lo - r, r - hi) * viol
            keep = vv >= FR * vv.max() if nviol else viol
            new = np.where(viol & keep, np.where(vlo, -1, 1), st)
            new[wrong] = 0
            if ends and len(ends) == len(wz):
                peel_cut(st, new, lam, x, mu0, Hd, aeq, Ain, lo, hi, C_, ends)
            elif it >= DAMP_AFTER and ends:
                for i in wz:
                    if i not in ends: new[i] = st[i]
        elif variant == 4:
            new[wrong] = 0
            if ends:
                peel_cut(st, new, lam, x, mu0, Hd, aeq, Ain, lo, hi, C_, ends)
        elif variant == 3:
            # two-phase but wrong rows in the interior of runs leave together with adds; run ends wait for feasibility
            if nviol == 0:
                new[wrong] = 0
                if ends and len(ends) == len(wz):
                    peel_cut(st, new, lam, x, mu0, Hd, aeq, Ain, lo, hi, C_, ends)
            else:
                for i in wz:
                    if i not in ends: new[i] = 0
        if np.array_equal(new, st):
            return x, st, it + 1
        st = new
    return None, st, maxit

if __name__ == '__main__':
    gait = sys.argv[1] if len(sys.argv) > 1 else 'trot'
    nsamp = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    variants = [int(v) for v in sys.argv[3].split(',')] if len(sys.argv) > 3 else [0, 1]
    if len(sys.argv) > 4:
        THETA = float(sys.argv[4])
    d = np.load(os.path.join(ROOT, 'gpurun_out', 'forma_dump_%s.npz' % gait))
    model, inst, ft, plan = d['model'], d['inst'], d['fs_timing'], d['fs_plan']
    C_, F = int(model["C"][0]), int(model["F"][0])
    kit = d['out']['iters']
    order = np.argsort(kit)[::-1]
    pick = list(order[:nsamp // 4]) + list(range(0, len(inst), max(1, len(inst) // (nsamp - nsamp // 4))))[:nsamp - nsamp // 4]
    res = {v: [] for v in variants}
    ref = {}
    t0 = time.time()
    for k in pick:
        Hd, g, A, lb, ub = T._dense(model, inst[k], ft, plan)
        for ax in range(2):
            q = T._axis(Hd, g, A, lb, ub, ax, C_, F)
            for v in variants:
                x, st, it = pdas(*q, C_, variant=v)
                if x is None: it = 999
                res[v].append(it)
                if v == variants[0]: ref[(k, ax)] = x
                elif x is not None and ref[(k, ax)] is not None:
                    assert np.abs(x - ref[(k, ax)]).max() < 1e-7, (k, ax, v)
    for v in variants:
        a = np.array(res[v])
        print('variant', v, 'n', len(a), 'mean %.2f p90 %d p99 %d max %d fails %d' % (a[a < 999].mean(), np.percentile(a, 90), np.percentile(a, 99), a[a<999].max(), (a == 999).sum()), np.bincount(np.minimum(a, 30)).tolist())
    print('time', time.time() - t0)
