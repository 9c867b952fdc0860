import sys, os, numpy as np
sys.path[:0]=['nl-partsol_b200','oracle','tests']
from nlps_b200 import engine
from util import load_points
for case in ('dp','mn'):
    z=load_points(case); X,Y=z['inputs'],z['outputs']
    r=engine.stress_points(2,str(z['mat_type']),z['mat_params'],float(z['tol_radial']),int(z['maxiter_radial']),X[:,0:5],X[:,5:10],X[:,10],X[:,11:16],X[:,16],X[:,17])
    got=np.concatenate([r['stress'],r['b_e_n1'],r['eps_n1'][:,None],r['kappa_n1'][:,None],r['W'][:,None],r['C_ep']],axis=1)
    np.save(f'gpurun_out/points_{case}.npy',got)
    s=np.maximum(np.abs(Y[:,0:5]).max(axis=1),1e-9); e=np.abs(got[:,0:5]-Y[:,0:5]).max(axis=1)/s
    print(case,'n',len(e),'bad>1e-10',int((e>1e-10).sum()),'max',e.max(),'status!=0',int((r['status']!=0).sum()))
