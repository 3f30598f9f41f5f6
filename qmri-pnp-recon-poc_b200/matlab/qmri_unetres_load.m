function net = qmri_unetres_load(weights, in_nc)
% weights: 1x64 cell of single arrays = UNetRes.state_dict() values in key order, each in PyTorch memory
% order (e.g. exported with scipy.io.savemat after .permute to reverse dims, or read from the .pt via Python).
% Replaces importONNXNetwork + assembleNetwork (main_recon_tsmis_FFT.m:138-152); the handle is what
% param.net receives instead of @(x) denoiseImage_PnP_ADMM(x, Net, true, false).
net = qmri_b200_mex('net_load', in_nc, weights);
end
