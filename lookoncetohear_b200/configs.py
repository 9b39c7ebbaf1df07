"""The two shape contracts of the reference (configs/tsh.json:5-19, configs/embed.json:5-10)."""
TSH_PARAMS = dict(embed_dim=256, stft_chunk_size=128, stft_pad_size=64, num_ch=2, D=64, L=4, I=1, J=1, B=3,
                  H=64, local_atten_len=50, use_attn=True, lookahead=True, chunk_causal=True)
EMBED_PARAMS = dict(embed_dim=256, num_ch=2, n_fft=128, stride=64, num_blocks=3)
