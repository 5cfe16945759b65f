"""Offline fit of the engine's per-round choice of inner-BnB kernel shape (engine.cu: run_inner_batch).  Input: the per-round
lines of GOICP_ROUND_STATS=1 for six golden runs, each forced into every shape (profiles/r2x_rounds/, captured with
GOICP_BNB_VARIANT=lat|thr|q5 scripts/golden_run.py NAME): the rounds of a run are the same in every shape (same results), so any
rule can be scored by adding up the measured kernel time of the shape it would have picked.  Output: profiles/r2x_shape_rule_fit.txt."""
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import re, numpy as np, sys
cfgs=['bunny_s0.1_mse1e-3','bunny_s0.1_mse5e-4','bunny_s0.033_mse1e-3','spanner_s0.02_mse3e-4','skull_s0.03_mse1e-3','spanner_s0.02_mse1e-4']
def load(f):
    rows=[]
    for l in open(f):
        m=re.match(r"\[round\] shape (\d+) forecast max (\d+) sum (\d+); \[round\] tasks (\d+) kernel ([\d.]+) ms; slowest task ([\d.]+) Mcyc \(pops (\d+), level (-?\d+)\); max pops (\d+); total pops (\d+); sum task cycles ([\d.]+) Mcyc",l)
        if m: rows.append([float(x) for x in m.groups()])
    return np.array(rows)
D={}
for c in cfgs:
    D[c]={v:load(os.path.join(ROOT, 'profiles', 'r2x_rounds', f'r2x_rounds_{c}_{v}.txt')) for v in ['lat','thr','q5']}
    n=[len(D[c][v]) for v in D[c]]
    assert len(set(n))==1,(c,n)
SM=148;cl=4;slots=SM//cl
def rule_forecast(r_lat, a_thr=(1.4,1.45), a_q=(2.0,2.1), margin_q=1.0, margin_thr=1.0, prev=None):
    # r_lat row: fields idx 1 = forecast max, 2 = forecast sum
    pm,ps=r_lat[1],r_lat[2]
    if ps<=0: return 'lat'
    t_lat=max(pm,ps/slots); t_thr=max(a_thr[0]*pm,a_thr[1]*ps/(2*slots)); t_q=max(a_q[0]*pm,a_q[1]*ps/(5*slots))
    best='lat'; tb=t_lat
    if t_thr*margin_thr<tb: best='thr'; tb=t_thr*margin_thr
    if t_q*margin_q<tb: best='q5'
    return best
def evaluate(rulefn):
    out={}
    for c in cfgs:
        d=D[c]; tot=0; picks={'lat':0,'thr':0,'q5':0}
        for i in range(len(d['lat'])):
            ch=rulefn(c,i); tot+=d[ch][i,4]; picks[ch]+=1
        out[c]=(round(tot,2),picks)
    return out
def show(name,res):
    print(name, {c:res[c][0] for c in cfgs})
for v in ['lat','thr','q5']:
    show('all-'+v, evaluate(lambda c,i: v))
show('oracle', evaluate(lambda c,i: min(['lat','thr','q5'], key=lambda v: D[c][v][i,4])))
# old rule: previous round measured (lat units) -> lat/thr
def old_rule(c,i):
    if i==0: return 'lat'
    r=D[c]['lat'][i-1]; maxc=r[5]; sumc=r[10]
    t_lat=max(maxc,sumc/slots); t_thr=max(1.43*maxc,1.54*sumc/(2*slots))
    return 'lat' if t_lat<=t_thr else 'thr'
show('old(prev measured)', evaluate(old_rule))
show('forecast m=1', evaluate(lambda c,i: rule_forecast(D[c]['lat'][i])))
for mq in [1.5,2,3]:
    show(f'forecast margin_q={mq} margin_thr={mq}', evaluate(lambda c,i: rule_forecast(D[c]['lat'][i],margin_q=mq,margin_thr=mq)))
# rule on number of tasks & forecast sum only (ignore forecast max): dense if sum/ (5*slots)*2.1 > K * typical?
def rule_sum(c,i,K):
    r=D[c]['lat'][i]; ps=r[2]; nt=r[3]
    if ps<=0: return 'lat'
    avg=ps/nt
    # throughput-bound if tasks*avg/slots >> expected max ~ K*avg
    return 'q5' if nt/ (5*slots) * 2.1 > K else ('thr' if nt/(2*slots)*1.45 > K else 'lat')
for K in [2,3,4,6,8]:
    show(f'tasks-only K={K}', evaluate(lambda c,i: rule_sum(c,i,K)))
# hybrid: previous round's measured max (in lat units, known from run) and forecast sum
def hybrid(c,i,ch_prev=[None]):
    r=D[c]['lat'][i]; ps=r[2]
    if i==0 or ps<=0: return 'lat'
    prev=D[c]['lat'][i-1]; ratio=prev[10]/max(prev[9],1)  # Mcyc per pop measured last round (lat units)
    pm=max(r[1],prev[8])  # forecast max vs last round's actual max pops
    t_lat=max(pm,ps/slots); t_thr=max(1.4*pm,1.45*ps/(2*slots)); t_q=max(2.0*pm,2.1*ps/(5*slots))
    return min([('lat',t_lat),('thr',t_thr),('q5',t_q)],key=lambda x:x[1])[0]
show('hybrid max(forecast,prev actual max)', evaluate(hybrid))
print('----')
def gen(alpha, beta, a_thr=(1.4,1.45), a_q=(2.0,2.1), bias=False):
    def f(c,i):
        r=D[c]['lat'][i]; ps=r[2]
        if i==0 or ps<=0: return 'lat'
        prev=D[c]['lat'][i-1]
        pm=max(beta*r[1],alpha*prev[8])
        if bias and prev[2]>0: ps=ps*min(4.0,max(0.25,prev[9]/prev[2]))
        t_lat=max(pm,ps/slots); t_thr=max(a_thr[0]*pm,a_thr[1]*ps/(2*slots)); t_q=max(a_q[0]*pm,a_q[1]*ps/(5*slots))
        return min([('lat',t_lat),('thr',t_thr),('q5',t_q)],key=lambda x:x[1])[0]
    return f
orc=evaluate(lambda c,i: min(['lat','thr','q5'], key=lambda v: D[c][v][i,4]))
def score(res): return sum(res[c][0]/orc[c][0] for c in cfgs)/len(cfgs)
best=[]
for alpha in [0,0.5,1,1.5,2]:
  for beta in [0,0.5,1,1.5,2,3]:
    for bias in [False,True]:
      for aq in [(2.0,2.1),(2.0,2.6),(2.4,2.1)]:
        res=evaluate(gen(alpha,beta,a_q=aq,bias=bias)); best.append((score(res),alpha,beta,bias,aq,{c:res[c][0] for c in cfgs}))
best.sort(key=lambda x:x[0])
for b in best[:8]: print(b)
print('old',score(evaluate(old_rule)),'forecast',score(evaluate(lambda c,i: rule_forecast(D[c]['lat'][i]))))
