function [K] = setup_subsampling_spiralgrided(N, M, S, V)
% Drop-in for main_files/subsampling_patterns/setup_subsampling_spiralgrided.m (same signature, same returned handles).
% K.for / K.adj are the reference's handles (P*x and P'*x, :41-42) backed by the GPU operator, so the closure lines
% main_recon_tsmis_FFT.m:228-229 run unedited; K.handle / K.size let qmri_fft_operator(K) build the fused F instead.
K.handle = qmri_b200_mex('op_spiral', N, M, S, double(real(V)));
K.size = [N, M, size(V, 2)];
h = K.handle; n = N * M * size(V, 2);
K.for = @(x) qmri_b200_mex('p_for', h, x);
K.adj = @(x) qmri_b200_mex('p_adj', h, x, n);
end
