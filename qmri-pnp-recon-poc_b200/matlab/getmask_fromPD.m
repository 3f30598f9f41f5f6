function mask = getmask_fromPD(PD, thresh)
% Drop-in for main_files/utils/getmask_fromPD.m (threshold of |PD|/max, hole filling, binarisation) on the GPU.
mask = qmri_b200_mex('mask', PD, thresh);
end
