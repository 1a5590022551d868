"""Model zoo: the five architectures the reference ships / trains."""
from .cnn import CNN_Net
from .resnet import resnet50, ResNet
from .inception import Inception3, inception_v3
from .gpt import GPT, GPTConfig
from .bert import BertForPreTraining, BertConfig, bert_pretraining_loss, to_hf_state_dict

__all__ = ['CNN_Net', 'resnet50', 'ResNet', 'Inception3', 'inception_v3', 'GPT', 'GPTConfig',
           'BertForPreTraining', 'BertConfig', 'bert_pretraining_loss', 'to_hf_state_dict']
