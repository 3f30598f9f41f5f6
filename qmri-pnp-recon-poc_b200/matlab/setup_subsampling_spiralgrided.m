function [K] = setup_subsampling_spiralgrided(N, M, S, V)
% Drop-in for main_files/subsampling_patterns/setup_subsampling_spiralgrided.m (same signature).
% Returns a struct with the operator handle; F.forward / F.adjoint are built by qmri_fft_operator(K).
% K.for / K.adj (the bare sparse matrix P) are not provided: the reference only uses them inside
% F.forward / F.adjoint (main_recon_tsmis_FFT.m:228-229), which run fused on the GPU.
K.handle = qmri_b200_mex('op_spiral', N, M, S, double(real(V)));
K.size = [N, M, size(V, 2)];
end
