"""One launch of the log-mel kernel on the bench shape, for ncu (tools/mel_profile.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_to_speech_b200.stft import TacotronSTFT

stft = TacotronSTFT()
x = (torch.randn(128, 1723 * 256, device="cuda") * 0.1).clamp_(-1, 1)
for _ in range(3):
    stft.mel_spectrogram(x)
torch.cuda.synchronize()
