function noise_map = build_noise_map(noise_std, rows, cols)
% Same as main_files/utils/build_noise_map.m.
noise_map = repmat(noise_std, rows, cols);
end
