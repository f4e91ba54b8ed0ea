import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audio_only_speech_separation_b200 import _lib, ops
B=16; S,K=82,100; P=B*S*K
dev=torch.device("cuda"); torch.manual_seed(0)
lstm=torch.nn.LSTM(64,128,1,batch_first=True,bidirectional=True).cuda()
pack=ops.LstmPack(lstm)
G0=torch.randn(P,1024,device=dev)*0.5
L=_lib.lib()
for mode in (0,2,3):
  _lib.check(L.dp_set_lstm_pipeline(mode))
  for save in (1,0):
    for useH in (1,0):
      G=torch.empty_like(G0); H=torch.empty(P,256,device=dev); C=torch.empty(P,256,device=dev)
      ts=[]
      for it in range(4):
        G.copy_(G0)
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.dp_lstm_recurrence_f32(_lib.ptr(pack.buf),_lib.ptr(G),_lib.ptr(H) if useH else None,_lib.ptr(C) if save else None,B*S,K,1<<30,0,K,1,save,0,_lib.stream_ptr()))
        e1.record(); torch.cuda.synchronize()
        if it>=1: ts.append(e0.elapsed_time(e1))
      print(json.dumps({"mode":mode,"save":save,"H":useH,"us":round(1e3*sum(ts)/len(ts),1)}),flush=True)
