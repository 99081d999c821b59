"""Constructor kwargs of the shipped configurations (hyper-parameter names and values are contract):
configs/mae/mae_HeadCT.yaml:31-50, configs/dino/dino_HeadCT.yaml:33-68, configs/downstream/vit_HeadCT_cq500.yaml:34-51,
mapped to kwargs as main_pretrain_mae.py:103-123 / main_pretrain_dino.py:110-166 / main_downstream.py:119-152 do."""

MAE_HEADCT = dict(input_size=96, patch_size=12, mask_ratio=0.75, in_chans=3, dropout_rate=0., spatial_dims=3,
                  patch_embed="conv", pos_embed="sincos", encoder_depth=12, encoder_embed_dim=768,
                  encoder_mlp_dim=3072, encoder_num_heads=12, decoder_depth=8, decoder_embed_dim=768,
                  decoder_mlp_dim=3072, decoder_num_heads=16, norm_pix_loss=False, use_bias=True)
MAE_TRAIN = dict(base_lr=1.5e-4, betas=(0.9, 0.95), weight_decay=0.05)

VIT_DINO = dict(in_chans=3, img_size=96, patch_size=12, hidden_size=768, mlp_dim=3072, num_layers=12, num_heads=12,
                patch_embed="conv", pos_embed="sincos", classification=False, dropout_rate=0., spatial_dims=3,
                num_register_tokens=4, qkv_bias=True)
DINO_HEAD = dict(in_dim=768, out_dim=65536, use_bn=False, norm_last_layer=True, nlayers=3, hidden_dim=2048,
                 bottleneck_dim=256)
DINO_LOSS = dict(out_dim=65536, ncrops=4, warmup_teacher_temp=0.04, teacher_temp=0.04, warmup_teacher_temp_epochs=30,
                 nepochs=200)
DINO_TRAIN = dict(base_lr=5e-4, betas=(0.9, 0.999), weight_decay=0.04, weight_decay_end=0.4, momentum_teacher=0.999)

VIT_DOWNSTREAM = dict(in_chans=3, img_size=96, patch_size=12, hidden_size=768, mlp_dim=3072, num_layers=12, num_heads=12,
                      patch_embed="conv", pos_embed="sincos", classification=False, dropout_rate=0., spatial_dims=3,
                      num_register_tokens=0, qkv_bias=True)
VIT_EXTRACT = dict(VIT_DOWNSTREAM, qkv_bias=False)       # released weights have no qkv bias (notebook cell 2-3)
