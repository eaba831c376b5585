-- Drop-in replacement for the reference's src/lua/npe.lua: `M = require 'npe'` returns the same `model` table
-- (create, fp, fp_batch, bp, update_position, update_angle, cuda/float/clearState, fields network / criterion /
-- identitycriterion / theta.params / theta.grad_params / neighbor_masks / mp) with every tensor operation executed by
-- libnpe_cuda.so.  Reads the same globals (mp, config_args) the reference reads.
require 'nn'
require 'NPECuda'
require 'IdentityCriterion'
local npeffi = require 'npe_ffi'
local C, ffi = npeffi.C, npeffi.ffi

local model = {}
model.__index = model

function model.create(mp_, preload, model_path)               -- npe.lua:72-102
    local self = setmetatable({}, model)
    self.mp = mp_
    assert(self.mp.input_dim == self.mp.object_dim * self.mp.num_past)
    assert(self.mp.out_dim == self.mp.object_dim * self.mp.num_future)
    self.criterion = nn.MSECriterion(false)                    -- kept for checkpoint compatibility (data_utils.lua:210-230)
    self.identitycriterion = nn.IdentityCriterion()
    self.network = nn.NPECuda(self.mp)
    if preload then
        local checkpoint = torch.load(model_path)
        local src = checkpoint.model.theta and checkpoint.model.theta.params or checkpoint.model.network.weight
        self.network.weight:copy(src)                           -- same flat layout as the reference's getParameters()
        self.network:push()
    end
    self.theta = {}
    self.theta.params, self.theta.grad_params = self.network:getParameters()
    collectgarbage()
    return self
end

function model:cuda() end
function model:float() end
function model:clearState() self.network:clearState() end

local function flat_target(y) return y:contiguous():view(y:size(1), -1) end

function model:fp(params_, batch, sim)                         -- npe.lua:225-247
    if params_ ~= self.theta.params then self.theta.params:copy(params_) end
    self.network:push()
    self.theta.grad_params:zero()
    local this, context, y = unpack(batch)
    self.network.target = flat_target(y)
    local prediction = self.network:forward({this, context})
    local l = self.network.loss3
    return l[1], prediction, l[2], l[3]
end

function model:fp_batch(params_, batch, sim)                   -- npe.lua:250-284
    if params_ ~= self.theta.params then self.theta.params:copy(params_) end
    self.network:push()
    local this, context, y = unpack(batch)
    local B = this:size(1)
    local yy = flat_target(y)
    npeffi.check(self.network.handle, C.npe_batch_upload(self.network.handle, npeffi.fptr(this:contiguous()),
                 npeffi.fptr(context:contiguous()), npeffi.fptr(yy), B, context:size(2)))
    local pred, le = torch.FloatTensor(B, 22), torch.FloatTensor(B, 3)
    npeffi.check(self.network.handle, C.npe_forward_per_example(self.network.handle, npeffi.fptr(pred), npeffi.fptr(le)))
    return le:select(2, 1):clone(), pred, le:select(2, 2):clone(), le:select(2, 3):clone()
end

function model:bp(batch, prediction, sim)                      -- npe.lua:290-336
    self.theta.grad_params:zero()
    self.network:accGradParameters(nil, nil, 1)
    return self.theta.grad_params
end

-- the mask of the last batch, as a table of (B) tensors like select_neighbors returns (npe.lua:170-206)
function model:pull_neighbor_masks()
    local h = self.network.handle
    local F, K = C.npe_num_foci(h), C.npe_ctx_slots(h)
    local mask = torch.ByteTensor(F, K)
    npeffi.check(h, C.npe_build_pairs(h, nil, ffi.cast('uint8_t*', torch.data(mask)), nil))
    self.neighbor_masks = {}
    for c = 1, K do table.insert(self.neighbor_masks, mask:select(2, c):float()) end
    return self.neighbor_masks
end

return model
