"""E2VID-shaped recurrent U-Net forward used as the CONSUMER of BASELINE config 5 (benchmark infrastructure, not product).

Random-initialised torch modules with the shapes of the reference's `E2VIDRecurrent` under config/train_v2v_e2vid_10k.yaml
(:18-30; model/model.py:194-223, model/unet.py:252-310, model/submodules.py:99-118,179-235):
head Conv(5->32,k5)+ReLU; 3 encoders Conv(k5,s2)+ReLU -> ConvLSTM(k3) with 32->64->128->256 channels; 2 residual blocks at
256; 3 decoders bilinear x2 -> Conv(k5)+ReLU 256->128->64->32; skip = sum; prediction Conv(32->1,k1); no normalisation.
It consumes the voxels where the simulator wrote them: `[B,T,5,Hp,Wp]` already in the /16-padded layout
(model/train_utils.py:322-326), one `[B,5,Hp,Wp]` slice per step, states reset per sequence.  cuDNN does the convolutions.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class ConvLSTM(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.gates = nn.Conv2d(2 * ch, 4 * ch, 3, padding=1)
        self.ch = ch

    def forward(self, x, state):
        if state is None:
            z = torch.zeros_like(x)
            state = (z, z)
        h, c = state
        i, f, o, g = self.gates(torch.cat((x, h), 1)).chunk(4, 1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        return h, (h, c)


class ResBlock(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.c1, self.c2 = nn.Conv2d(ch, ch, 3, padding=1), nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return F.relu(self.c2(F.relu(self.c1(x))) + x)


class E2VIDShaped(nn.Module):
    def __init__(self, num_bins=5, base=32, num_encoders=3, num_res=2):
        super().__init__()
        self.head = nn.Conv2d(num_bins, base, 5, padding=2)
        chans = [base * 2 ** i for i in range(num_encoders + 1)]
        self.enc = nn.ModuleList(nn.Conv2d(chans[i], chans[i + 1], 5, stride=2, padding=2) for i in range(num_encoders))
        self.lstm = nn.ModuleList(ConvLSTM(chans[i + 1]) for i in range(num_encoders))
        self.res = nn.ModuleList(ResBlock(chans[-1]) for _ in range(num_res))
        self.dec = nn.ModuleList(nn.Conv2d(chans[i + 1], chans[i], 5, padding=2) for i in reversed(range(num_encoders)))
        self.pred = nn.Conv2d(base, 1, 1)
        self.states = [None] * num_encoders

    def reset_states(self):
        self.states = [None] * len(self.enc)

    def forward(self, x):
        x = F.relu(self.head(x))
        head, blocks = x, []
        for i, (conv, cell) in enumerate(zip(self.enc, self.lstm)):
            x, self.states[i] = cell(F.relu(conv(x)), self.states[i])
            blocks.append(x)
        for r in self.res:
            x = r(x)
        for i, conv in enumerate(self.dec):
            x = F.relu(conv(F.interpolate(x + blocks[-1 - i], scale_factor=2, mode="bilinear", align_corners=False)))
        return self.pred(x + head)

    @torch.no_grad()
    def forward_sequence(self, padded_events):
        """padded_events [B,T,5,Hp,Wp] -> [B,T,1,Hp,Wp] (model/train_utils.py:309-348 without the pad copy)."""
        self.reset_states()
        return torch.stack([self.forward(padded_events[:, t]) for t in range(padded_events.shape[1])], dim=1)
