mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sr_gpu.py -m gpu -q --timeout 300 > gpurun_out/r10_sr.log 2>&1; grep -n "AssertionError\|^E   " gpurun_out/r10_sr.log | head -20
timeout 300 python - <<'PY'
import numpy as np, torch
from oracle import sr as S
from visiondepth3d_b200 import merged_pipeline as MP
from visiondepth3d_b200.synth import synth_frame
for nc,w,h in ((4,96,54),(16,160,90),(32,320,180),(32,157,93)):
    sd=S.srvgg_state_dict(num_conv=nc, seed=3)
    eng=MP.SrEngine({k:v.numpy() for k,v in sd.items()})
    fr,_=synth_frame(2,w,h,"natural")
    out=eng.upscale(fr)
    with torch.no_grad():
        rf=S.srvgg_forward(sd, torch.from_numpy(S.preprocess_esr(fr))).numpy()
    ref=S.postprocess_esr(rf)
    d=np.abs(out.astype(int)-ref.astype(int))
    print(nc,w,h,'max',d.max(),'frac>0',(d>0).mean(),'frac>1',(d>1).mean(),'mean',d.mean(), 'residual rms', float(np.sqrt(((rf[0]-np.repeat(np.repeat(S.preprocess_esr(fr)[0],4,1),4,2))**2).mean())))
    eng.close()
PY
