import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench_configs
from audiotokenization_b200 import configs, synth
from audiotokenization_b200.model import BigCodecModel
from audiotokenization_b200.vq import module as M
cfg = configs.get_config("base")
enc_sd, dec_sd = synth.make_state_dicts(cfg, seed=0)
model = BigCodecModel(cfg, enc_sd, dec_sd, device="cuda", precision="bf16x3")
for chunk in (128, 64, 128, 64, 48, 128, 64, 48):
    M.LSTM_WAVEFRONT_CHUNK[0] = chunk
    line = bench_configs.round_trip(model, 64, 10.0, 8, check=False)
    print(chunk, round(line["ms_per_step"], 2), round(line["audio_s_per_s"]))
