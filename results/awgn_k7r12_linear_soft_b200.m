% B200 sweep (tools/sweep_awgn.py --code k7r12 --labels linear --ebn0 0:6:1 --frames 200000 --min-frame-errors 100), L=1024, 1 GPU
% format follows the reference results/awgn_channel.m
snr_k7r12_linear_soft = [0 1 2 3 4 5 6];
bits_k7r12_linear_soft = [204800000 204800000 204800000 204800000 204800000 204800000 1024000000];
ber_k7r12_linear_soft = [0.1538086670 0.0402762256 0.0053703125 0.0004435059 0.0000277100 0.0000019385 0.0000001738];
