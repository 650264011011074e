"""CPU test of the ALGORITHM the subgraph-augmentation kernel implements (csrc/augment.cu): a Python restatement of its explicit
container emulation -- networkx insertion orders as arrays, CPython's set of small ints as an open-addressing table -- against
oracle/subgraph.py, which runs on Python's own dicts and sets (and is pinned by vectors from the reference's functions).  The GPU
test (tests/test_gpu_augment.py) then checks the kernel itself against the oracle."""
import numpy as np

from oracle import subgraph as osg
from molclr_b200.synth import random_molecule

class PySet:
    def __init__(s): s.mask=7; s.fill=0; s.slot=[-1]*8
    @staticmethod
    def insert_clean(table, mask, key):
        perturb=key; i=key&mask
        while True:
            if table[i]<0: table[i]=key; return
            if i+9<=mask:
                for j in range(1,10):
                    if table[i+j]<0: table[i+j]=key; return
            perturb>>=5; i=(i*5+1+perturb)&mask
    def add(s,key):
        perturb=key; i=key&s.mask
        while True:
            probes=9 if i+9<=s.mask else 0
            e=i; found=False
            while True:
                if s.slot[e]<0: found=True; break
                if s.slot[e]==key: return
                e+=1
                if probes==0: break
                probes-=1
            if found: break
            perturb>>=5; i=(i*5+1+perturb)&s.mask
        s.slot[e]=key; s.fill+=1
        if s.fill*5 < s.mask*3: return
        newsize=8
        while newsize <= s.fill*4: newsize<<=1
        old=[k for k in s.slot if k>=0]
        s.mask=newsize-1; s.slot=[-1]*newsize
        for k in old: PySet.insert_clean(s.slot,s.mask,k)
    def items(s): return [k for k in s.slot if k>=0]

def emulate(n, bonds, center, percent, mode):
    rank=[-1]*n; nodes=[]; adj=[[] for _ in range(n)]
    for s,e in bonds:
        s=int(s); e=int(e)
        if rank[s]<0: rank[s]=len(nodes); nodes.append(s)
        if rank[e]<0: rank[e]=len(nodes); nodes.append(e)
        if e not in adj[s]:
            adj[s].append(e)
            if e!=s: adj[e].append(s)
    gl=[]
    for a in range(n):
        l=[]
        for r in range(max(rank[a],0)):
            y=nodes[r]
            if y in adj[a]: l.append(y)
        for y in adj[a]:
            if rank[y]>=rank[a]: l.append(y)
        gl.append(l)
    removed=[False]*n; nrem=0
    num=int(np.floor(len(nodes)*percent))
    if num>0 and rank[center]>=0:
        temp=[center]
        while nrem<num and temp:
            st=PySet()
            for u in temp:
                for v in gl[u]:
                    if removed[v]: continue
                    if v not in temp: st.add(v)
            for t in temp:
                if nrem<num: removed[t]=True; nrem+=1
            temp=st.items()
    keep=[]
    for s,e in bonds:
        ok = not removed[int(s)] and not removed[int(e)]
        if ok and mode==1: ok = rank[int(s)]<rank[int(e)]
        keep.append(ok)
    return removed, keep

def test_container_emulation_matches_oracle_on_random_molecules():
  rng=np.random.default_rng(0)
  bad=0; total=0
  for trial in range(500):
      x,bonds,battr=random_molecule(rng, mean_atoms=rng.choice([8,25,45,70]), std_atoms=8)
      if trial%5==0 and len(bonds)>3:   # shuffle bond order and orientation: exercises insertion orders
          perm=rng.permutation(len(bonds)); bonds=bonds[perm]; battr=battr[perm]
          flip=rng.random(len(bonds))<0.5; bonds=np.where(flip[:,None], bonds[:,::-1], bonds)
      n=len(x)
      for center in rng.choice(n, size=min(3,n), replace=False):
          for mode,percent in ((1,0.25),(2,float(rng.uniform(0,0.2)))):
              total+=1
              rem,keep=emulate(n,bonds,int(center),percent,mode)
              if mode==1:
                  xv,ei,ea,removed=osg.subgraph_view(x,bonds,battr,int(center),percent)
                  want_keep=[]
                  g,_=osg.remove_subgraph(osg.build_graph(bonds), int(center), percent)
                  ge=osg.edge_list(g)
                  want_keep=[(int(s),int(e)) in ge for s,e in bonds]
              else:
                  g,removed=osg.remove_subgraph(osg.build_graph(bonds), int(center), percent, stop_when_exhausted=True)
                  ge=osg.edge_list(g)
                  want_keep=[((int(s),int(e)) in ge) or ((int(e),int(s)) in ge) for s,e in bonds]
              if sorted(removed)!=[i for i,r in enumerate(rem) if r] or want_keep!=keep:
                  bad+=1
                  if bad<5: print('MISMATCH',trial,center,mode,percent,sorted(removed),[i for i,r in enumerate(rem) if r])
  assert total > 2000 and bad == 0, (total, bad)
