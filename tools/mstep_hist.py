import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cpelnano_b200 as cpn
eng=cpn.Engine(0); lib=eng.lib
nb=cpn.simulations.make_nano_batch(3000,M=30,seed=9)
rng=np.random.default_rng(1)
phi0=cpn.simulations.gen_phi0s(rng,nb.R,1)[:,0]
ES,_=lib.estep_batch(nb.grp_off,nb.Nl,nb.rho_l,nb.d_l,nb.read_off,nb.log_pyx_u,nb.log_pyx_m,phi0,with_logpy=False)
sol,conv,cnt=lib.mstep_batch(nb.grp_off,nb.Nl,nb.rho_l,nb.d_l,np.diff(nb.read_off),ES,phi0)
print("first M-step from random starts: converged frac",conv.mean()," iters hist",np.bincount(cnt[:,0],minlength=21))
# second-iteration-like: start from clamped solution
phi1=np.where(conv[:,None],np.clip(sol,[-20,-150,-50],[20,150,50]),phi0)
ES,_=lib.estep_batch(nb.grp_off,nb.Nl,nb.rho_l,nb.d_l,nb.read_off,nb.log_pyx_u,nb.log_pyx_m,phi1,with_logpy=False)
sol2,conv2,cnt2=lib.mstep_batch(nb.grp_off,nb.Nl,nb.rho_l,nb.d_l,np.diff(nb.read_off),ES,phi1)
print("second M-step: converged frac",conv2.mean()," iters hist",np.bincount(cnt2[:,0],minlength=21))
phi=phi1
for k in range(3,9):
    phi=np.where(conv2[:,None],np.clip(sol2,[-20,-150,-50],[20,150,50]),phi)
    ES,_=lib.estep_batch(nb.grp_off,nb.Nl,nb.rho_l,nb.d_l,nb.read_off,nb.log_pyx_u,nb.log_pyx_m,phi,with_logpy=False)
    sol2,conv2,cnt2=lib.mstep_batch(nb.grp_off,nb.Nl,nb.rho_l,nb.d_l,np.diff(nb.read_off),ES,phi)
    print(f"M-step {k}: converged frac",conv2.mean()," iters hist",np.bincount(cnt2[:,0],minlength=21))
