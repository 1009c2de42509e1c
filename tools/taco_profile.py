"""A short B200 Tacotron2 decode (16 x 86 tokens, 64 frames, direct launches) for ncu launch lists."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_to_speech_b200.tacotron2 import Tacotron2, Tacotron2HParams, generate_tacotron2_weights
from text_to_speech_b200.tts import synthetic_texts

hp = Tacotron2HParams()
w = generate_tacotron2_weights(hp, 77)
w["decoder/gate_output/bias"][:] = -10.0
m = Tacotron2(hp, w, device="cuda", b200_lstm_weights=os.environ.get("TACO_LSTM", "split_bf16"))
toks = np.stack(synthetic_texts(16, 99, 86, 86))
frames = int(os.environ.get("TACO_FRAMES", "64"))
chunk = int(os.environ.get("TACO_CHUNK", "0"))
for _ in range(2):
    m.infer(toks, max_length=frames, early_stopping=False, decoder="b200", graph_chunk=chunk, return_attention=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
m.infer(toks, max_length=frames, early_stopping=False, decoder="b200", graph_chunk=chunk, return_attention=False)
e1.record()
torch.cuda.synchronize()
print(f"infer {frames} frames, chunk {chunk}: {e0.elapsed_time(e1):.2f} ms")
