function [x] = PnP_ADMM(y, param)
% Drop-in for main_files/algorithms/PnP_ADMM/PnP_ADMM.m (same signature and param fields:
% iter, gamma, cg_tol, F, gt_tsmi, X0, net, denoiser_type, noise_map).
% param.F must come from qmri_fft_operator; param.net is either a handle returned by
% qmri_unetres_load (whole loop on the GPU) or any function handle @(v) (pluggable prox, host hop).
if ~isfield(param, 'denoiser_type'), param.denoiser_type = 'single_level'; end
x = qmri_b200_mex('pnp_admm', param.F.handle, double(y), param);
end
