% k7r12, block length 1024, ref labels, 8 B200 GPU(s), 0.13 s, vdec_sweep (ncclAllReduce of the error counters per batch)
n_gpus_k7r12_ref_soft = 8;
snr_k7r12_ref_soft = [0 2 4 6 8];
bits_k7r12_ref_soft = [8192000000 8192000000 8192000000 8192000000 8192000000];
frame_errors_k7r12_ref_soft = [8000000 8000000 7999996 7999671 7999354];
ber_k7r12_ref_soft = [2.128608e-01 1.036330e-01 4.117184e-02 2.392591e-02 2.117155e-02];
