"""Restatement of the ``complexnn`` module that ``DCCRN/DCCRN_cprs.py:6`` imports but the
reference tree does NOT contain (un-vendored, un-pinned; provenance: upstream
huyanxin/DeepComplexCRN ``complexnn.py``).  TEST INFRASTRUCTURE.

Written from the published algorithm:
  ComplexConv2d          real = conv_r(x_r) - conv_i(x_i); imag = conv_i(x_r) + conv_r(x_i);
                         symmetric padding[0] on the first spatial axis inside the conv, causal
                         left padding of padding[1] frames on the time axis before it;
  ComplexConvTranspose2d same combination with nn.ConvTranspose2d;
  NavieComplexLSTM       real = L_r(x_r) - L_i(x_i); imag = L_r(x_i) + L_i(x_r); optional
                         r_trans / i_trans projection;
  complex_cat            concatenate the real halves, then the imaginary halves.
The shipped checkpoints pin the STRUCTURE (key names and shapes load strictly into these
classes); no reference test pins the numerics => "parity unpinned" at this boundary, the
restatement is the de-facto specification.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class ComplexConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=(1, 1), stride=(1, 1), padding=(0, 0), dilation=1,
                 groups=1, causal=True, complex_axis=1):
        super().__init__()
        self.padding, self.causal, self.complex_axis = padding, causal, complex_axis
        self.real_conv = nn.Conv2d(in_channels // 2, out_channels // 2, kernel_size, stride,
                                   padding=[padding[0], 0], dilation=dilation, groups=groups)
        self.imag_conv = nn.Conv2d(in_channels // 2, out_channels // 2, kernel_size, stride,
                                   padding=[padding[0], 0], dilation=dilation, groups=groups)

    def forward(self, x):
        if self.padding[1] != 0 and self.causal:
            x = F.pad(x, [self.padding[1], 0, 0, 0])
        else:
            x = F.pad(x, [self.padding[1], self.padding[1], 0, 0])
        r, i = torch.chunk(x, 2, self.complex_axis)
        real = self.real_conv(r) - self.imag_conv(i)
        imag = self.imag_conv(r) + self.real_conv(i)
        return torch.cat([real, imag], self.complex_axis)


class ComplexConvTranspose2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=(1, 1), stride=(1, 1), padding=(0, 0),
                 output_padding=(0, 0), causal=False, complex_axis=1, groups=1):
        super().__init__()
        self.complex_axis = complex_axis
        self.real_conv = nn.ConvTranspose2d(in_channels // 2, out_channels // 2, kernel_size, stride,
                                            padding=padding, output_padding=output_padding, groups=groups)
        self.imag_conv = nn.ConvTranspose2d(in_channels // 2, out_channels // 2, kernel_size, stride,
                                            padding=padding, output_padding=output_padding, groups=groups)

    def forward(self, x):
        r, i = torch.chunk(x, 2, self.complex_axis)
        real = self.real_conv(r) - self.imag_conv(i)
        imag = self.imag_conv(r) + self.real_conv(i)
        return torch.cat([real, imag], self.complex_axis)


class NavieComplexLSTM(nn.Module):
    def __init__(self, input_size, hidden_size, projection_dim=None, bidirectional=False, batch_first=False):
        super().__init__()
        self.input_dim, self.rnn_units = input_size // 2, hidden_size // 2
        self.real_lstm = nn.LSTM(self.input_dim, self.rnn_units, num_layers=1, bidirectional=bidirectional,
                                 batch_first=False)
        self.imag_lstm = nn.LSTM(self.input_dim, self.rnn_units, num_layers=1, bidirectional=bidirectional,
                                 batch_first=False)
        fac = 2 if bidirectional else 1
        if projection_dim is not None:
            self.projection_dim = projection_dim // 2
            self.r_trans = nn.Linear(self.rnn_units * fac, self.projection_dim)
            self.i_trans = nn.Linear(self.rnn_units * fac, self.projection_dim)
        else:
            self.projection_dim = None

    def forward(self, inputs):
        real, imag = inputs
        r2r = self.real_lstm(real)[0]
        r2i = self.imag_lstm(real)[0]
        i2r = self.real_lstm(imag)[0]
        i2i = self.imag_lstm(imag)[0]
        real_out = r2r - i2i
        imag_out = i2r + r2i
        if self.projection_dim is not None:
            real_out = self.r_trans(real_out)
            imag_out = self.i_trans(imag_out)
        return [real_out, imag_out]

    def flatten_parameters(self):
        self.real_lstm.flatten_parameters()
        self.imag_lstm.flatten_parameters()


def complex_cat(inputs, axis):
    real, imag = [], []
    for data in inputs:
        r, i = torch.chunk(data, 2, axis)
        real.append(r)
        imag.append(i)
    return torch.cat([torch.cat(real, axis), torch.cat(imag, axis)], axis)


class ComplexBatchNorm(nn.Module):       # never instantiated: every script passes use_cbn=False
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("use_cbn=True is not used by any decode script")
