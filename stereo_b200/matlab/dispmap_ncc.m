% dispmap_ncc -- B200 build of the NCC data-term model (surface of dispmap_ncc.m:5-277).
% The NCC volume, its WTA initialisation and the parabola-interpolated sampling run on the GPU
% (sb_ncc_volume / sb_ncc_best_disp / sb_ncc_sample through sb_builders_mex).
% One optional trailing constructor argument extends the reference: patchsize (default 2, the
% value the reference hard-codes at dispmap_ncc.m:24).
classdef dispmap_ncc < dispmap_super
	properties (SetAccess = protected)
		ncc;
		disparities;
		unary_weight;
		smooth;
		tol;
	end
	methods
		function self = dispmap_ncc(images, disparities, kernel, unary_weight, tol, patchsize)
			self = self@dispmap_super(images, kernel);
			if nargin < 6, patchsize = 2; end
			self.disparities = disparities;
			self.unary_weight = unary_weight;
			self.smoothness_kernel = kernel;
			self.tol = tol;
			self.ncc = sb_builders_mex('ncc_volume', double(images{1}), double(images{2}), double(disparities(:)), patchsize);
			init_solution(self);
		end
		function set.tol(self, tol)
			if (tol < 0), error('Tolerance weight must be positive'); end
			self.tol = tol;
			update_energy(self);
		end
		function set.unary_weight(self, weight)
			if (weight < 0), error('Unary weight must be positive'); end
			self.unary_weight = weight;
			update_energy(self);
		end
		function proposal = generate_new_plane_RANSAC(self, x, y, r)
			% Plane through the WTA disparities within radius r of (x, y), fitted and replicated on the GPU
			% (sb_plane_from_disparity; fit_plane_to_points below is the host form of the same fit).
			[~, proposal] = sb_builders_mex('plane_from_disparity', best_disp_from_ncc(self), x, y, r, self.smoothness_kernel);
		end
		function p = fit_plane_to_points(self, points)
			n = size(points, 2);
			A = -(points - repmat(mean(points, 2), [1 n]))';
			p = zeros(4, 1);
			if (self.smoothness_kernel == 1)
				w = ones(n, 1);
				for irls_iteration = 1:20 %#ok<NASGU>
					[~, ~, V] = svd(repmat(w, [1 3]) .* A, 'econ');
					p(1:3) = V(:, end);
					w = sqrt(abs(A * V(:, end)));
				end
			else
				[~, ~, V] = svd(A, 'econ');
				p(1:3) = V(:, end);
			end
			p(4) = -(p(1:3)' * mean(points(1:3, :), 2));
			p = p / p(3);
		end
		function display(self)
			fprintf('Disparity map with normalized cross correlation unary term \n');
			fprintf('Disparity levels %d in range [%g, %g]. \n', numel(self.disparities), max(self.disparities), min(self.disparities));
			display@dispmap_super(self);
			fprintf('Unary weight         : %g \n', self.unary_weight);
			fprintf('Tolerance            : %g \n', self.tol);
		end
		function restart(self)
			init_solution(self);
		end
	end
	methods (Access = protected)
		function U = unary_cost(self, assignment)
			if isempty(self.ncc), U = zeros(prod(self.sz), 1); return; end
			disps = disparitymap_from_assignment(self, assignment);
			U = sb_builders_mex('ncc_sample', self.ncc, double(self.disparities(:)), disps(:), self.unary_weight, 1);
			U = U(:);
		end
		function init_solution(self)
			best_disp = best_disp_from_ncc(self);
			a = zeros(4, prod(self.sz));
			a(3, :) = 1;
			a(4, :) = -best_disp(:);
			self.assignment = a;
		end
		function best_disp = best_disp_from_ncc(self)
			best_disp = sb_builders_mex('ncc_best_disp', self.ncc, double(self.disparities(:)));
		end
		function nccs = sample_ncc_from_disp(self, new_pixel_disps)
			nccs = sb_builders_mex('ncc_sample', self.ncc, double(self.disparities(:)), new_pixel_disps(:), 1, 0);
		end
	end
end
