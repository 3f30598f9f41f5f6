function Y = qmri_awgn(Y, snr, seed)
% GPU replacement for Y = awgn(Y, snr, 'measured') (main_recon_tsmis_FFT.m:243; Communications Toolbox): complex white noise
% of power mean(|Y|^2)/10^(snr/10) per column, Philox counter generator keyed by seed (reproducible; not MATLAB's stream).
if nargin < 3, seed = 0; end
Y = qmri_b200_mex('awgn', complex(Y), snr, seed);
end
