import numpy as np, torch, sys
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from helpers import *
from gym_genesis_b200.model import franka_cube_pick_model
from gym_genesis_b200.sim import BatchedSim
from oracle.binding import Oracle
m=franka_cube_pick_model(); o=Oracle(m)
n=256
rng=np.random.default_rng(11)
Q,QD,T=random_states(m,rng,n,cube_on_ground=True); QA=np.zeros_like(QD)
sim=BatchedSim(m,n)
sim.set_state(Q,QD,QA); sim.step(torch.as_tensor(f32(T)))
q,qd,qa=(t.cpu().numpy().astype(np.float64) for t in sim.get_state())
qo,qdo,qao=Q.copy(),QD.copy(),QA.copy()
info,_=o.step(qo,qdo,qao,T,iterations=100,tol=1e-12)
sc=np.array([np.abs(o.dynamics(Q[k],QD[k],T[k])["a_smooth"]).max()*0.01 for k in range(n)])
eq,ev=substep_errors(Q,QD,qo,qdo,q,qd,sc)
print("eq",eq.max(),"ev",ev.max())
k=int(np.argmax(ev)); print(k, info[k], "qd err", (qd-qdo)[k], "dqd ref", (qdo-QD)[k])
print("qacc gpu", qa[k]); print("qacc ref", qao[k])
st=sim.stats().cpu().numpy(); print(st)
np.set_printoptions(precision=4, suppress=True, linewidth=200)
print("q err by column", np.abs(q-qo).max(0))
print("qd err by column", np.abs(qd-qdo).max(0))
print("rows with big err", np.where(np.abs(q-qo).max(1)>1e-3)[0][:40])
