function m = qmri_recon_metrics(qmap, qmap0, foreground_mask, X, X0)
% The metrics block of main_recon_tsmis_FFT.m:328-384 on the GPU; fields named like the script's variables.
v = qmri_b200_mex('metrics', qmap, double(qmap0), double(foreground_mask), X, X0);
names = {'tsmi_mean_psnr','tsmi_mean_ssim','t1_mae','t1_psnr','t1_ssim','t2_mae','t2_psnr','t2_ssim','pd_mae','pd_psnr','pd_ssim'};
for i = 1:11, m.(names{i}) = v(i); end
end
