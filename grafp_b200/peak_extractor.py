"""Peak point-cloud front end (reference: peak_extractor.py), the caller feeding GraphEncoder.

On CUDA fp32 spectrograms the whole forward (min-max normalisation, position ramps, strided convolution, ReLU,
flattening) is one kernel that writes the point cloud as node rows - the layout the encoder's stem consumes - and its
backward one kernel + a reduction; anything else runs the PyTorch ops of the reference.
"""
import torch
import torch.nn as nn

from . import ops


class GPUPeakExtractorv2(nn.Module):
    """min-max normalise the spectrogram, append time / frequency ramps, 7x7 conv with stride
    (peak_stride, 1) + ReLU, flatten to a (B, n_filters, n_mels * n_frames / peak_stride) point cloud
    (reference: peak_extractor.py:11-82)."""

    def __init__(self, cfg):
        super().__init__()
        self.blur_kernel = cfg['blur_kernel']
        self.n_filters = cfg['n_filters']
        self.stride = cfg['peak_stride']
        self.convs = nn.Sequential(
            nn.Conv2d(3, self.n_filters, kernel_size=self.blur_kernel, stride=(self.stride, 1),
                      padding=(self.blur_kernel[0] // 2, self.blur_kernel[1] // 2)),
            nn.ReLU(),
        )
        self._ramps = {}
        self.init_weights()

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def _position_ramps(self, n_mels, n_frames, device):
        # the reference pre-sizes these to bsz_train // n_gpus and rebuilds them on a shape mismatch
        # (peak_extractor.py:32-41,71-77); the values only depend on (n_mels, n_frames)
        key = (n_mels, n_frames, device)
        if key not in self._ramps:
            t = torch.linspace(0, 1, steps=n_frames, device=device).view(1, 1, 1, n_frames).expand(1, 1, n_mels, n_frames)
            f = torch.linspace(0, 1, steps=n_mels, device=device).view(1, 1, n_mels, 1).expand(1, 1, n_mels, n_frames)
            self._ramps[key] = torch.cat((t, f), dim=1)
        return self._ramps[key]

    def forward(self, spec_tensor):
        conv = self.convs[0]
        if ops.peak_extract_supported(spec_tensor, conv):
            # (B, F, N, 1) channels-last -> the reference's (B, F, N) view of the same memory (node rows).  Under
            # torch.autocast the kernel still computes and returns fp32 (cuDNN would run this 3 -> 8 channel 7x7
            # convolution through conv2d_grouped_direct_kernel at 0.5 ms per view); the stem's convolution casts it.
            return ops.peak_extract(spec_tensor, conv).squeeze(-1)
        lo = torch.amin(spec_tensor, dim=(1, 2), keepdim=True)
        hi = torch.amax(spec_tensor, dim=(1, 2), keepdim=True)
        peaks = ((spec_tensor - lo) / (hi - lo)).unsqueeze(1)
        B, _, n_mels, n_frames = peaks.shape
        ramps = self._position_ramps(n_mels, n_frames, peaks.device).expand(B, 2, n_mels, n_frames)
        feature = self.convs(torch.cat((ramps, peaks), dim=1))
        return feature.reshape(B, feature.shape[1], -1)
