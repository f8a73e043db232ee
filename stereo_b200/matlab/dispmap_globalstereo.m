% dispmap_globalstereo -- B200 build of the photo-consistency model of Woodford et al.
% (surface of dispmap_globalstereo.m:10-480).  The unary (plane -> disparity -> warp -> bilinear
% gather -> ephoto) runs on the GPU (sb_photo_unary); segmentation, SegPln proposals and the
% 3-D display remain host-side preprocessing exactly as in the reference and are reached through
% the imrender package when it is on the path.
classdef dispmap_globalstereo < dispmap_super
	properties (SetAccess = public)
		tol;
	end
	properties (SetAccess = protected)
		options;
		ephoto;
		P;
		disp_range;
		disparity_factor;
		disps;
		d_min;
		d_step;
		start_disparity;
	end
	methods
		function self = dispmap_globalstereo(images, P, disp_range, disparity_factor, options)
			self = self@dispmap_super(images, options.smoothness_kernel);
			self.tol = options.disp_thresh;
			if max(abs(P([1:6 9]) - [1 0 0 0 1 0 1])) > 1e-12
				error('First image must be reference image');
			end
			self.P = permute(P(:,:,:), [2 1 3]);
			self.disp_range = disp_range;
			self.disparity_factor = disparity_factor;
			self.disps = sort(disp_range(1)*disparity_factor : disp_range(2)*disparity_factor, 'descend');
			self.d_min = self.disps(end);
			self.d_step = self.disps(1) - self.d_min;
			self.dnorm = [self.d_min self.d_step];
			self.options = options;
			preprocess(self);
			self.start_disparity = rand(self.sz) * self.d_step + self.d_min;
			init_solution(self);
		end
	end
	methods
		function corr = segpln_wta(self)
			% The winner-takes-all window-matching disparity map SegPln fits its planes to, on the GPU.
			ims = cellfun(@double, self.images, 'UniformOutput', false);
			corr = sb_builders_mex('segpln_wta', cat(4, ims{:}), permute(self.P(:,:,:), [2 1 3]), self.disps, ...
				self.options.window, self.options.col_thresh, 0.07);
		end
	end
	methods (Access = protected)
		function init_solution(self)
			a = zeros(4, prod(self.sz));
			a(3, :) = 1;
			a(4, :) = -self.start_disparity(:);
			self.assignment = a;
		end
		function U = unary_cost(self, assignment)
			if isempty(self.P), U = zeros(prod(self.sz), 1); return; end
			U = sb_builders_mex('photo_unary', double(self.images{1}), double(self.images{2}), self.P(:,:,2), ...
				assignment, self.d_min, self.d_step, self.options.col_thresh);
		end
		function preprocess(self)
			% Segmentation-weighted smoothness: lambda_h inside a mean-shift segment, lambda_l across.
			ref = uint8(self.images{1});
			colors = size(ref, 3);
			if colors == 1, ref = repmat(ref, [1 1 3]); end
			self.options.planar = 0;
			segment = vgg_segment_ms(ref, self.options.seg_params(1), self.options.seg_params(2), self.options.seg_params(3));
			self.improve = (self.options.improve > 0);
			EW = sb_builders_mex('smooth_weights', uint32(segment), self.options.lambda_h, self.options.lambda_l, ...
				numel(self.images) / ((self.options.connect == 8) + 1));
			self.ephoto = @(F) log(2) - log(exp(sum(F .^ 2, 2) * (-1 / (self.options.col_thresh * colors))) + 1);
			self.smooth_weights = EW;
			if (self.smoothness_kernel == 2)
				self.smooth_weights = self.smooth_weights / self.tol;
				self.tol = self.tol^2;
			end
		end
	end
end
