import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from oracle import cases
from helpers import solver_from_refcase, rel_l2
for name, mk, nst in (("boxper n=6", lambda: cases.case_boxper((3,3,3),6), 3), ("3dboxper", cases.case_3dboxper, 10), ("3dboxpec", cases.case_3dboxpec, 10),
                      ("dielectric-nocb", None, 5)):
    if mk is None:
        c = cases.case_3ddielectric(True); c.s.userinc = type(c.s.userinc)()  # drop callback
    else:
        c = mk()
    s = solver_from_refcase(c)
    # stage-level check
    c.stage(1); s.stage(1); s.synchronize()
    print(name, 'stage1 relL2 H', rel_l2(s.hn, c.hn), 'E', rel_l2(s.en, c.en))
    for rk in range(2,6): c.stage(rk); s.stage(rk)
    c.s.time += c.s.dt; s.set_time(c.s.time, c.s.dt)
    c.step(nst); s.step(nst)
    print(name, 'after', nst+1, 'steps relL2 H', rel_l2(s.hn, c.hn), 'E', rel_l2(s.en, c.en), 'time', s.time, c.time)
    if c.ifpml: print('  pml B', rel_l2(s.get_array('pmlbn'), c.pmlbn), 'D', rel_l2(s.get_array('pmldn'), c.pmldn))
    ms, nl = s.last_step_ms(); print('  ms', ms, 'launches', nl)
    shn, sen = c.usersol(c, c.time)
    l2, linf = s.cem_error(shn, sen); l2o, linfo = c.errors(c.usersol)
    print('  l2 gpu', l2.max(), 'oracle', max(l2o), 'linf', linf.max(), max(linfo))
    s.close()
