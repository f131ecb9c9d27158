from _fallthrough import extend as _extend
_extend(__path__, "layers")
