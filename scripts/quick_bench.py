import sys, time, importlib, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
pkg = importlib.import_module("cuda-go-icp_b200")
G='/root/repo/tests/golden/'
model = np.fromfile(G+'bunny_model_s0.1_seed1234.f32', np.float32).reshape(-1,3)
data = np.fromfile(G+'bunny_data_s0.1_seed1235.f32', np.float32).reshape(-1,3)
for mode in (0,1):
    g = pkg.GoICP(1e-3); g.pModel, g.pData = model, data; g.dt_mode = mode
    t=time.time(); g.BuildDT(); t1=time.time(); print('dt mode', mode, 'build s', t1-t)
    t=time.time(); g.BuildDT(); t1=time.time(); print('dt mode', mode, 'build s (2nd)', t1-t)
    if mode == 0: grid, meta = g.GetDT()
    g.close()
for mse in (1e-3, 7e-4, 5e-4):
  for spec in (0, 4, 36):
    g = pkg.GoICP(mse); g.pModel, g.pData = model, data; g.spec_cubes = spec
    g.SetDT(grid, meta)
    g.Register()
    r = g.result
    print('mse', mse, 'spec', spec, {k: (round(v,5) if isinstance(v,float) else v) for k,v in r.items() if k not in ('R','t')})
    g.close()
